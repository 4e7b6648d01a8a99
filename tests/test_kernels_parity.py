"""The kernels (through the C ABI) against the oracle and the reference's golden vectors.

Every test runs twice: on the host-emulation build of the kernel sources (CPU tier) and, marked
``gpu``, on the nvcc build for sm_100a (the product).  Integer / index work is compared bit for bit.
K-Means: the kernel must equal oracle/kmeans_restated.py (scikit-learn's arithmetic restated, itself equal
to every recorded scikit-learn call) label for label, so the pipeline equals the reference's own output.
"""
import numpy as np
import pytest
import torch

from conftest import golden, injected_kmeans, same_component
from oracle import rhccq_oracle as O
from roibasedimagecompression_b200 import ops, pipeline
from roibasedimagecompression_b200.encoder import compression as C
from roibasedimagecompression_b200.synth import synth, tile_regions


# --------------------------------------------------------------------------- a1
def test_unique_colors_golden(backend):
    g = golden("unique_colors.npz")
    for i in range(3):
        r = C.get_all_unique_colors(g[f"crop{i}"], (7 * i, 3 * i), as_arrays=True)
        assert np.array_equal(r["palette"], g[f"palette{i}"])
        assert np.array_equal(r["indices"], g[f"indices{i}"])
        assert r["actual_colors"] == len(g[f"palette{i}"]) and r["top_left"] == (7 * i, 3 * i)


def test_unique_colors_edge_cases(backend):
    assert C.get_all_unique_colors(None, (0, 0)) is None                       # clustering.py:9-10
    assert C.get_all_unique_colors(np.zeros((0, 4, 3), np.uint8), (0, 0)) is None
    one = np.full((1, 1, 3), 7, np.uint8)
    r = C.get_all_unique_colors(one, (3, 4))
    assert r["palette"] == [[7, 7, 7]] and r["indices"] == [0]
    black = np.zeros((5, 6, 3), np.uint8)
    r = C.get_all_unique_colors(black, (0, 0), as_arrays=True)
    assert r["palette"].tolist() == [[0, 0, 0]] and not r["indices"].any()
    rng = np.random.default_rng(3)
    wide = rng.integers(0, 256, size=(37, 113, 3)).astype(np.uint8)             # every pixel its own colour, ragged size
    r = C.get_all_unique_colors(wide, (0, 0), as_arrays=True)
    ref = O.get_all_unique_colors(wide, (0, 0))
    assert np.array_equal(r["palette"], ref["palette"]) and np.array_equal(r["indices"], ref["indices"])
    # thousands of colours in one key-ordered bucket (one R, a narrow band of G) while R and G span their whole
    # range: the hashed path hands the segment to the sorting path
    skew = np.stack([np.full(4096, 10), rng.integers(0, 32, 4096), rng.integers(0, 256, 4096)], axis=1).astype(np.uint8)
    skew[0], skew[1] = (0, 255, 7), (255, 0, 9)
    skew = skew.reshape(64, 64, 3)
    r = C.get_all_unique_colors(skew, (0, 0), as_arrays=True)
    ref = O.get_all_unique_colors(skew, (0, 0))
    assert np.array_equal(r["palette"], ref["palette"]) and np.array_equal(r["indices"], ref["indices"])
    # a smooth segment: few values of R, many colours per value (large buckets that are still ranked by counting)
    yy, xx = np.mgrid[0:64, 0:64]
    smooth = np.stack([100 + xx // 16, 50 + (yy * 3 + rng.integers(0, 3, (64, 64))) % 200, rng.integers(0, 256, (64, 64))], axis=2).astype(np.uint8)
    r = C.get_all_unique_colors(smooth, (0, 0), as_arrays=True)
    ref = O.get_all_unique_colors(smooth, (0, 0))
    assert np.array_equal(r["palette"], ref["palette"]) and np.array_equal(r["indices"], ref["indices"])


# --------------------------------------------------------------------------- a2
def test_cluster_params_device_matches_python(backend):
    be = backend
    ns = np.array([0, 1, 2, 7, 99, 100, 3164, 9999, 65535, 250000], dtype=np.int32)
    for q in (1.0, 10.0, 20.0, 33.0, 40.0, 60.0, 99.0, 100.0):
        dev = ops.cluster_params(be, torch.from_numpy(ns).to(be.device),
                                 torch.full((len(ns),), q, dtype=torch.float64, device=be.device))
        want = [O.compute_clustering_params(int(n), q)[2] for n in ns]
        assert dev.cpu().tolist() == want, q


# --------------------------------------------------------------------------- a3'
def test_palette_dbscan_matches_sklearn_golden(backend):
    g = golden("dbscan_palette.npz")
    for ti in range(4):
        pal = g[f"pal{ti}"]
        for q in g["qs"]:
            eps, _, _ = O.compute_clustering_params(len(pal), int(q))
            lab = C.clustering.dbscan_palette_labels(pal, eps)
            assert np.array_equal(lab, g[f"lab{ti}_q{q}"]), (ti, int(q))


def test_palette_dbscan_ties_and_tiny(backend):
    # eps^2 integer -> pairs at d2 == eps^2 are float64 ties; q = 75 -> eps 32, q = 50 -> eps 64
    rng = np.random.default_rng(5)
    base = rng.integers(1, 200, size=(40, 3))
    pts = [base]
    for d in ([32, 0, 0], [0, 32, 0], [0, 0, 32], [64, 0, 0], [0, 64, 0]):
        pts.append(np.clip(base + d, 1, 255))
    pal = np.unique(np.concatenate(pts), axis=0).astype(np.uint8)
    for q in (50, 75, 87.5):
        eps = 128 - 1.28 * q
        want = O._dbscan_sklearn(pal, eps, 1) if _have_sklearn() else O.dbscan_labels(pal, eps, 1, colour_scale=True)
        assert np.array_equal(C.clustering.dbscan_palette_labels(pal, eps), want), q
    for n in (1, 2, 3):
        p = np.array([[10, 10, 10], [10, 10, 12], [200, 10, 10]], np.uint8)[:n]
        assert np.array_equal(C.clustering.dbscan_palette_labels(p, 1.0),
                              O.dbscan_labels(p, 1.0, 1, colour_scale=True))


def _have_sklearn():
    try:
        import sklearn  # noqa: F401
        return True
    except Exception:
        return False


# --------------------------------------------------------------------------- a3 / a4
def _golden_cluster_cases(g):
    for c in range(int(g["n_cases"])):
        q = int(g[f"q{c}"])
        comp = {"palette": g[f"in_palette{c}"], "indices": g[f"in_indices{c}"],
                "shape": tuple(int(v) for v in g[f"shape{c}"]), "top_left": (0, 0)}
        eps, _, m = O.compute_clustering_params(len(comp["palette"]), q, "lab")
        yield c, q, comp, eps, m


def test_cluster_palette_equals_restated_oracle(backend):
    """Kernel K-Means == oracle/kmeans_restated.py bit for bit, so palette and indices are identical."""
    g = golden("cluster_palette.npz")
    for c, q, comp, eps, m in _golden_cluster_cases(g):
        want = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m)
        got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                as_arrays=True)
        assert np.array_equal(got["palette"], want["palette"]), (c, q)
        assert np.array_equal(got["indices"], want["indices"]), (c, q)
        assert got["compressed_colors"] == want["compressed_colors"]
        assert got["original_unique_colors"] == want["original_unique_colors"]


def test_cluster_palette_equals_reference_when_sklearn_assignment_injected(backend):
    """With scikit-learn's own K-Means labels injected, the result equals the reference's golden output."""
    g = golden("cluster_palette.npz")
    km = injected_kmeans(g)
    for c, q, comp, eps, m in _golden_cluster_cases(g):
        ref = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                kmeans_impl=km)
        assert np.array_equal(ref["palette"], g[f"out_palette{c}"])
        # the oracle's LUT (old row -> new row) is the leaf assignment to inject
        pal_in = comp["palette"]
        lut = np.zeros(len(pal_in), np.int32)
        lut[np.asarray(comp["indices"])] = ref["indices"]            # every row is used by some pixel
        used = np.zeros(len(pal_in), bool)
        used[np.asarray(comp["indices"])] = True
        assert used.all()
        got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                as_arrays=True, leaf_override=lut)
        assert np.array_equal(got["palette"], g[f"out_palette{c}"]), (c, q)
        assert np.array_equal(got["indices"], g[f"out_indices{c}"]), (c, q)


def test_cluster_palette_with_duplicate_rows(backend):
    """A palette passed on unmerged may hold a colour twice (two clusters truncating to one mean,
    merging.py:16-21): the K-Means sees both rows, find_color_index (:803-808) maps only the first."""
    rng = np.random.default_rng(9)
    for trial in range(6):
        base = np.unique(np.clip(rng.integers(60, 200, 3) + rng.integers(-12, 13, (int(rng.integers(30, 300)), 3)), 1, 255)
                         .astype(np.uint8), axis=0)
        pal = np.concatenate([base, base[rng.integers(0, len(base), int(rng.integers(1, 6)))]])
        pal = pal[rng.permutation(len(pal))]
        comp = {"palette": pal, "indices": rng.integers(0, len(pal), 64), "shape": (8, 8), "top_left": (0, 0)}
        q = float(rng.choice([10, 20, 40]))
        eps, _, m = O.compute_clustering_params(len(pal), q)
        want = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m)
        got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                as_arrays=True)
        assert np.array_equal(got["palette"], want["palette"]), trial
        assert np.array_equal(got["indices"], want["indices"]), trial


def test_pipeline_refuses_duplicate_rows_passed_on_unmerged(backend):
    """One tile per class: stage 1's palette reaches stage 2 unmerged; when it holds a colour twice the entry
    tables cannot follow the reference (first row only) and the pipeline must refuse, not approximate."""
    from roibasedimagecompression_b200._lib import RhccqError
    rng = np.random.default_rng(4)
    refused = agreed = 0
    for trial in range(40):
        img = np.clip(rng.integers(1, 6, (8, 16, 1)) + rng.integers(0, 2, (8, 16, 3)), 1, 255).astype(np.uint8)   # near-grey
        roi, non = tile_regions(8, 16, 8)
        try:
            pal, idx = _encode_device(backend, img, roi, non)
        except RhccqError as e:
            assert "holds a colour twice" in str(e)
            refused += 1
            continue
        want = O.encode_image(img, roi, non)
        assert np.array_equal(pal, want["palette"]) and np.array_equal(idx, want["indices"]), trial
        agreed += 1
    assert agreed > 0


def test_cluster_palette_degenerate(backend):
    allblack = {"palette": [[0, 0, 0]], "indices": [0, 0, 0, 0], "shape": (2, 2), "top_left": (0, 0)}
    assert C.cluster_palette_colors_parallel(20, allblack, eps=102.4, min_samples=1) is allblack   # clustering.py:197-199
    two = {"palette": [[0, 0, 0], [5, 6, 7]], "indices": [0, 1, 1, 0], "shape": (2, 2), "top_left": (1, 2)}
    r = C.cluster_palette_colors_parallel(20, two, eps=102.4, min_samples=1, max_colors_per_cluster=1)
    w = O.cluster_palette_colors_parallel(20, two, eps=102.4, min_samples=1, max_colors_per_cluster=1)
    assert r["palette"] == w["palette"].tolist() and r["indices"] == w["indices"].tolist()
    with pytest.raises(NotImplementedError):
        C.cluster_palette_colors_parallel(20, two, eps=102.4, min_samples=2)


# --------------------------------------------------------------------------- a5
def test_merge_canvas_golden(backend):
    g = golden("merge_canvas.npz")
    comps = [{"top_left": tuple(int(v) for v in g[f"top_left{i}"]), "shape": tuple(int(v) for v in g[f"shape{i}"]),
              "palette": g[f"palette{i}"], "indices": g[f"indices{i}"]} for i in range(int(g["n_comp"]))]
    r = C.merge_region_components_simple(comps, tuple(int(v) for v in g["bbox"]), as_arrays=True)[0]
    assert np.array_equal(r["palette"], g["out_palette"])
    assert np.array_equal(r["indices"], g["out_indices"])
    assert tuple(r["shape"]) == tuple(g["out_shape"])
    assert r["actual_colors"] == len(g["out_palette"])


def test_merge_degenerate(backend):
    assert C.merge_region_components_simple([], (0, 0, 4, 4)) == []             # merging.py:13-14
    one = {"top_left": (0, 0), "shape": (1, 2), "palette": [[1, 2, 3]], "indices": [0, 0]}
    r = C.merge_region_components_simple([one], (0, 0, 4, 4))                   # :16-21
    assert r[0]["actual_colors"] == 1 and r[0]["palette"] == one["palette"] and r[0] is not one


def test_merge_components_with_empty_palettes(backend):
    """merging.py:52-82 with fewer than two components carrying palette rows: still a fresh canvas, palette =
    black + the colours the paint sequence meets, unused rows dropped."""
    rng = np.random.default_rng(31)
    full = {"top_left": (1, 2), "shape": (4, 5), "palette": np.array([[0, 0, 0], [9, 9, 9], [200, 3, 3], [7, 7, 7]], np.uint8),
            "indices": rng.integers(0, 3, 20)}                                   # row 3 is never used
    empty = {"top_left": (0, 0), "shape": (2, 2), "palette": np.zeros((0, 3), np.uint8), "indices": np.zeros(4, int)}
    for comps in ([full, empty], [empty, full], [empty, full, empty], [empty, empty]):
        want = O.merge_region_components_simple(comps, (0, 0, 8, 8))[0]
        got = C.merge_region_components_simple(comps, (0, 0, 8, 8), as_arrays=True)[0]
        assert np.array_equal(np.asarray(got["palette"]).reshape(-1, 3), want["palette"])
        assert np.array_equal(got["indices"], want["indices"])
        assert got["actual_colors"] == want["actual_colors"]


def test_merge_random_overlaps(backend):
    rng = np.random.default_rng(21)
    for trial in range(4):
        comps = []
        for _ in range(int(rng.integers(2, 7))):
            h, w = int(rng.integers(1, 12)), int(rng.integers(1, 12))
            m = int(rng.integers(1, 6))
            pal = rng.integers(0, 4, size=(m, 3)).astype(np.uint8) * 60          # few colours -> shared between comps
            pal = np.unique(pal, axis=0)
            comps.append({"top_left": (int(rng.integers(-3, 14)), int(rng.integers(-3, 14))), "shape": (h, w),
                          "palette": pal, "indices": rng.integers(0, len(pal), size=h * w)})
        want = O.merge_region_components_simple(comps, (0, 0, 16, 16))[0]
        got = C.merge_region_components_simple(comps, (0, 0, 16, 16), as_arrays=True)[0]
        assert np.array_equal(got["palette"], want["palette"]), trial
        assert np.array_equal(got["indices"], want["indices"]), trial


def test_merge_large_component_palettes(backend):
    """Components that own thousands of colours each (the rank of a colour is then found by sorting all colours
    of the group rather than by counting inside its component), next to many small ones."""
    rng = np.random.default_rng(5)
    comps = []
    for h, w, ncol in ((60, 70, 3000), (50, 64, 2500), (8, 8, 40), (70, 60, 1500), (5, 9, 20)):
        pal = np.unique(rng.integers(0, 32, size=(ncol, 3)).astype(np.uint8) * 8, axis=0)
        comps.append({"top_left": (int(rng.integers(0, 20)), int(rng.integers(0, 20))), "shape": (h, w),
                      "palette": pal, "indices": rng.integers(0, len(pal), size=h * w)})
    want = O.merge_region_components_simple(comps, (0, 0, 96, 96))[0]
    got = C.merge_region_components_simple(comps, (0, 0, 96, 96), as_arrays=True)[0]
    assert np.array_equal(got["palette"], want["palette"])
    assert np.array_equal(got["indices"], want["indices"])


# --------------------------------------------------------------------------- a6 / a7: the three stages
def _encode_device(be, img, roi, non, qualities=(20, 10)):
    H, W, _ = img.shape
    tab, lab = pipeline.table_from_regions((H, W), [roi, non], qualities)
    res = pipeline.encode_batch(be, torch.from_numpy(img[None].copy()).to(be.device),
                                torch.from_numpy(lab).to(be.device), tab)
    pipeline.finish_checks(res)
    return res.palette(0), res.index_image(0).reshape(-1).astype(np.int64)


def _golden_image(g):
    if "image" in g.files:
        return g["image"]
    import os
    from PIL import Image
    from conftest import GOLDEN
    return np.array(Image.open(os.path.join(GOLDEN, "Lenna.png")).convert("RGB"))


@pytest.mark.parametrize("name", ["pipeline_small.npz", "pipeline_synth.npz", "pipeline_lenna.npz"])
def test_pipeline_equals_reference_output(backend, name):
    """The three-stage encode against the palette and index plane the reference itself produced
    (tests/golden/make_golden.py ran its modules, scikit-learn K-Means included): bit for bit.
    pipeline_lenna is BASELINE.json configs[0] — Lenna 512x512 at rhccq_20_10."""
    g = golden(name)
    img = _golden_image(g)
    roi, non = tile_regions(img.shape[0], img.shape[1], int(g["tile"]))
    pal, idx = _encode_device(backend, img, roi, non)
    assert np.array_equal(pal, g["palette"])
    assert np.array_equal(idx, g["indices"].astype(np.int64))
    rec = pal[idx.reshape(img.shape[:2])]
    assert abs(O.psnr(rec, img) - float(g["psnr"])) < 1e-9


def _recorded_calls():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import kmeans_replay
    return [c for c in kmeans_replay.collect() if c[1] == "km" and len(c[2]) >= 3]


def _ranks(lab):
    return np.unique(np.asarray(lab), return_inverse=True)[1]


def test_kmeans_kernel_reproduces_recorded_sklearn_calls(backend):
    """rhccq_palette_split with max_cpc = -k is KMeans(k, random_state=42, n_init='auto').fit_predict: every
    call scikit-learn performed for the reference's goldens (3 693; every 4th on the emulation build), label
    for label."""
    calls = _recorded_calls()
    if backend.device.type != "cuda":
        calls = calls[::4]
    for tag, _, col, k, lab in calls:
        assert np.array_equal(ops.kmeans_labels(backend, col, k), _ranks(lab)), (tag, len(col), k)


def _kmeans_cases(rng, count):
    for t in range(count):
        mode = t % 7
        if mode == 0:
            col = rng.integers(1, int(rng.integers(3, 30)), (int(rng.integers(3, 40)), 3))
        elif mode == 1:
            col = np.clip(rng.integers(40, 200, 3) + rng.normal(size=(int(rng.integers(20, 1200)), 3)) * rng.uniform(1, 20), 1, 255)
        elif mode == 2:
            col = np.clip(rng.integers(40, 200, 3) + rng.normal(size=(int(rng.integers(1000, 4500)), 3)) * rng.uniform(5, 30), 1, 255)
        elif mode == 3:
            col = rng.integers(1, 48, (int(rng.integers(300, 1500)), 3))
        elif mode == 4:
            a = rng.integers(1, 256, (4, 3))
            n = int(rng.integers(5, 200))
            col = np.clip(a[rng.integers(0, 4, n)] + rng.integers(-2, 3, (n, 3)), 1, 255)
        elif mode == 5:
            col = rng.integers(100, 104, (int(rng.integers(3, 25)), 3))
        else:                          # duplicate colours: clusters go empty and are relocated
            base = rng.integers(1, 255, (int(rng.integers(2, 8)), 3))
            col = base[rng.integers(0, len(base), int(rng.integers(6, 80)))]
        col = col.astype(np.uint8)
        if mode != 6:
            col = np.unique(col, axis=0)
        n = len(col)
        if n < 3:
            continue
        if mode == 3:
            k = int(rng.integers(129, max(130, n // 2)))          # centre tables in global memory; k > 192: blocked dgemm
        elif mode in (0, 5):
            k = int(rng.integers(2, n + 1))
        elif mode == 2:
            k = int(rng.integers(2, 60))
        else:
            k = int(rng.integers(2, min(n, 12) + 1))
        yield col, min(k, n)


def test_kmeans_kernel_equals_oracle_on_random_palettes(backend):
    """Warp-level and CTA-level K-Means, shared and global centre tables, ties, k up to n, relocation."""
    from oracle import kmeans_restated as K
    rng = np.random.default_rng(17)
    for col, k in _kmeans_cases(rng, 140):
        assert np.array_equal(ops.kmeans_labels(backend, col, k), _ranks(K.kmeans_labels(col, k))), (len(col), k)


def test_kmeans_kernel_with_every_decision_in_float64(emu_exact_backend):
    """-DRHCCQ_KM_FORCE_EXACT: the integer / margin forms are bypassed and every decision (candidate draw,
    best candidate, nearest centre, tolerance) is taken on the sequential float64 evaluation — the same labels."""
    from oracle import kmeans_restated as K
    be = emu_exact_backend
    rng = np.random.default_rng(18)
    for col, k in _kmeans_cases(rng, 70):
        assert np.array_equal(ops.kmeans_labels(be, col, k), _ranks(K.kmeans_labels(col, k))), (len(col), k)
    for tag, _, col, k, lab in _recorded_calls()[::16]:
        assert np.array_equal(ops.kmeans_labels(be, col, k), _ranks(lab)), (tag, len(col), k)


def test_pipeline_irregular_segments_with_black(backend):
    """Non-rectangular segments, true-black pixels inside segments (subregions.py:395-421), a region
    with one segment (:679) and ROI/non-ROI overlap in the 3 px buffer (encoder/ROI/roi.py:685-718)."""
    H, W = 72, 96
    img = synth(H, W, 77, sigma=2.0)
    img[10:14, 20:30] = 0
    img[50, 60] = 0
    yy, xx = np.mgrid[0:H, 0:W]
    roi_mask = ((yy - 36) ** 2 + (xx - 48) ** 2) < 28 ** 2
    non_mask = ((yy - 36) ** 2 + (xx - 48) ** 2) >= 25 ** 2                   # overlaps the ROI ring
    seg = (1 + (yy // 24) * 4 + (xx + yy // 3) // 24).astype(np.int32)

    def region(mask):
        rows = np.flatnonzero(mask.any(axis=1)); cols = np.flatnonzero(mask.any(axis=0))
        r0, r1, c0, c1 = rows[0], rows[-1] + 1, cols[0], cols[-1] + 1
        bm = mask[r0:r1, c0:c1]
        return {"bbox": (int(r0), int(c0), int(r1), int(c1)), "bbox_mask": bm,
                "segments": np.where(bm, seg[r0:r1, c0:c1], 0).astype(np.int32)}
    small = np.zeros((H, W), bool)
    small[2:9, 80:93] = True
    roi = [region(roi_mask)]
    single = region(small)
    single["segments"] = np.where(single["bbox_mask"], 1, 0).astype(np.int32)
    non = [region(non_mask & ~small), single]
    want = O.encode_image(img, roi, non)
    pal, idx = _encode_device(backend, img, roi, non)
    assert np.array_equal(pal, want["palette"])
    assert np.array_equal(idx, want["indices"])


def test_stage_functions_equal_oracle(backend):
    """subregion_quantization -> region_quantization -> quantize_image through the reference-named shims."""
    g = golden("pipeline_small.npz")
    img = g["image"]
    H, W, _ = img.shape
    roi, non = tile_regions(H, W, int(g["tile"]))
    s1r = C.subregion_quantization(img, roi, 20, "ROI")
    s1n = C.subregion_quantization(img, non, 10, "nonROI")
    w1r = O.subregion_quantization(img, roi, 20, "ROI")
    w1n = O.subregion_quantization(img, non, 10, "nonROI")
    for got, want in ((s1r, w1r), (s1n, w1n)):
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert len(a) == len(b) and all(same_component(x, y) for x, y in zip(a, b))
    r2 = C.region_quantization(s1r, H, W, 40)
    n2 = C.region_quantization(s1n, H, W, 20)
    assert same_component(r2[0], O.region_quantization(w1r, H, W, 40)[0])
    assert same_component(n2[0], O.region_quantization(w1n, H, W, 20)[0])
    fin = C.quantize_image(r2 + n2, H, W, 60)
    want = O.encode_image(img, roi, non)
    assert same_component(fin, want)
    assert isinstance(fin["indices"], list) and isinstance(fin["palette"], list)   # the reference's list contract
    assert fin["indices_dtype"] == want["indices_dtype"]


def test_batch_of_images_equals_per_image(backend):
    be = backend
    H, W, tile, B = 64, 96, 32, 3
    imgs = np.stack([synth(H, W, 1234 + i) for i in range(B)])
    tab, lab = pipeline.table_from_tiles(B, H, W, tile)
    labels = np.repeat(lab, B, axis=1)
    res = pipeline.encode_batch(be, torch.from_numpy(imgs).to(be.device), torch.from_numpy(labels).to(be.device), tab)
    pipeline.finish_checks(res)
    roi, non = tile_regions(H, W, tile)
    for b in range(B):
        want = O.encode_image(imgs[b], roi, non)
        assert np.array_equal(res.palette(b), want["palette"]), b
        assert np.array_equal(res.index_image(b).reshape(-1).astype(np.int64), want["indices"]), b


# --------------------------------------------------------------------------- split kernel stress
def _split_case(rng, kind):
    if kind == "clumps":            # several well separated clumps: several large clusters at high quality
        centres = rng.integers(20, 236, size=(int(rng.integers(2, 7)), 3))
        pts = np.concatenate([c + rng.integers(-9, 10, size=(int(rng.integers(5, 400)), 3)) for c in centres])
    elif kind == "dups":            # few distinct values per channel: many exact distance ties, empty clusters
        pts = rng.integers(0, 6, size=(int(rng.integers(30, 600)), 3)) * int(rng.integers(1, 40)) + 1
    elif kind == "line":
        t = rng.integers(1, 256, size=int(rng.integers(3, 1500)))
        pts = np.stack([t, t, np.clip(t + rng.integers(-1, 2, size=t.size), 1, 255)], axis=1)
    else:
        pts = rng.integers(1, 256, size=(int(rng.integers(1, 2500)), 3))
    return np.unique(np.clip(pts, 1, 255).astype(np.uint8), axis=0)


@pytest.mark.parametrize("kind", ["clumps", "dups", "line", "uniform"])
def test_split_kernel_stress_equals_oracle(backend, kind):
    """Many shapes of palettes through DBSCAN + recursive K-Means split: warp-level and CTA-level splits,
    several roots, ties, empty-cluster relocation, k from 2 to n."""
    rng = np.random.default_rng({"clumps": 1, "dups": 2, "line": 3, "uniform": 4}[kind])
    for trial in range(10):
        pal = _split_case(rng, kind)
        n = len(pal)
        q = float(rng.choice([1, 5, 10, 20, 40, 60, 80, 90, 95, 99, 100]))
        comp = {"palette": pal, "indices": np.arange(n), "shape": (1, n), "top_left": (0, 0)}
        eps, _, m = O.compute_clustering_params(n, q)
        if trial % 3 == 2:
            m = int(rng.integers(1, max(2, n // 2)))
        want = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m)
        got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                as_arrays=True)
        assert np.array_equal(got["palette"], want["palette"]), (kind, trial, n, q, m)
        assert np.array_equal(got["indices"], want["indices"]), (kind, trial, n, q, m)


def test_pipeline_random_images_equal_oracle(backend):
    """A slice of tools/pipeline_fuzz.py in the suite: random small images, tile sizes and quality pairs through the
    three stages, palette and every index equal to the oracle's (the duplicate-rows refusal of DESIGN.md section 2
    is the only accepted alternative)."""
    from roibasedimagecompression_b200._lib import RhccqError
    rng = np.random.default_rng(606)
    equal = 0
    for c in range(40 if backend.device.type == "cuda" else 8):
        H, W = int(rng.integers(3, 8)) * 16, int(rng.integers(3, 10)) * 16
        tile = int(rng.choice([16, 32, 48, 64]))
        img = synth(H, W, int(rng.integers(0, 10 ** 6)), sigma=float(rng.choice([1.0, 3.0, 6.0])))
        if rng.random() < 0.3:
            img = np.clip((img // int(rng.integers(2, 9))) * int(rng.integers(1, 5)), 0, 255).astype(np.uint8)
        quals = (int(rng.choice([10, 20, 30, 50])), int(rng.choice([5, 10, 20])))
        roi, non = tile_regions(H, W, tile)
        want = O.encode_image(img, roi, non, roi_quality=quals[0], nonroi_quality=quals[1])
        try:
            pal, idx = _encode_device(backend, img, roi, non, quals)
        except RhccqError as e:
            assert "holds a colour twice" in str(e)
            continue
        assert np.array_equal(pal, want["palette"]), (c, H, W, tile, quals)
        assert np.array_equal(idx, np.asarray(want["indices"]).reshape(-1)), (c, H, W, tile, quals)
        equal += 1
    assert equal >= 6


# --------------------------------------------------------------------------- >= 10 000 colours
def _big_palette(seed, n):
    img = synth(256, 256, seed, sigma=3.0)
    pal = np.unique(img.reshape(-1, 3), axis=0)
    assert len(pal) >= n
    return pal[np.sort(np.random.default_rng(seed).choice(len(pal), n, replace=False))]


@pytest.mark.parametrize("n,q,black", [(10000, 10, False), (12000, 40, True), (10003, 20, False)])
def test_minibatch_branch_equals_restated_oracle(backend, n, q, black):
    """clustering.py:207-218: MiniBatchKMeans for >= 10 000 non-black colours, then the usual split of
    clusters larger than max_colors_per_cluster.  Kernel == oracle/minibatch_restated.py bit for bit."""
    pal = _big_palette(11 + n, n)
    if black:
        pal = np.concatenate([np.zeros((1, 3), np.uint8), pal])
    comp = {"palette": pal, "indices": np.arange(len(pal)), "shape": (1, len(pal)), "top_left": (0, 0)}
    eps, _, m = O.compute_clustering_params(len(pal), q)
    want = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m)
    got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m, as_arrays=True)
    assert np.array_equal(got["palette"], want["palette"])
    assert np.array_equal(got["indices"], want["indices"])


def test_minibatch_branch_equals_reference_output(backend):
    """The reference itself (tests/golden/make_golden.py minibatch: its cluster_palette_colors_parallel with
    scikit-learn's MiniBatchKMeans and KMeans underneath) on a 10 500-colour palette at two qualities: the
    kernels reproduce its palette, order included, and its indices — nothing injected."""
    g = golden("minibatch_palette.npz")
    pal = g["in_palette"]
    for c in range(int(g["n_cases"])):
        q = int(g[f"q{c}"])
        comp = {"palette": pal, "indices": np.arange(len(pal)), "shape": (1, len(pal)), "top_left": (0, 0)}
        eps, _, m = O.compute_clustering_params(len(pal), q, "lab")
        got = C.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m, as_arrays=True)
        assert np.array_equal(got["palette"], g[f"out_palette{c}"]), c
        assert np.array_equal(got["indices"], g[f"out_indices{c}"]), c


def test_minibatch_kernel_equals_sklearn_labels(backend):
    """rhccq_palette_minibatch on one palette is MiniBatchKMeans(...).fit_predict: labels recorded from scikit-learn
    itself (tests/golden/make_minibatch_sklearn.py) on five palettes of 10 000 - 15 739 colours, k = 126 ... 630;
    the oracle reproduces them too (it is what the other tests of this branch compare the kernel with)."""
    from oracle import minibatch_restated as MB
    g = golden("minibatch_sklearn.npz")
    n_cases = int(g["n_cases"])
    assert n_cases >= 4
    for c in range(n_cases if backend.device.type == "cuda" else 2):
        col, q, k = g[f"colors{c}"], float(g[f"q{c}"]), int(g[f"k{c}"])
        want = g[f"labels{c}"].astype(np.int64)
        lab, k_dev = ops.minibatch_labels(backend, col, q)
        assert k_dev == k
        assert np.array_equal(lab, want), (c, len(col), k, float((lab == want).mean()))
        if c < 2:
            assert np.array_equal(MB.minibatch_labels(col, k), want)


def test_pipeline_with_large_segments_minibatch_in_stage1(backend):
    """Segments of more than 10 000 colours (natural 128 px tiles hit this): stage 1 itself takes the
    MiniBatchKMeans branch, and the index plane needs no more than 16 bits."""
    H, W, tile = 128, 256, 128
    img = synth(H, W, 99, sigma=6.0)
    roi, non = tile_regions(H, W, tile)
    n_colours = len(np.unique(img[:, :128].reshape(-1, 3), axis=0))
    assert n_colours >= 10000
    want = O.encode_image(img, roi, non)
    pal, idx = _encode_device(backend, img, roi, non)
    assert np.array_equal(pal, want["palette"])
    assert np.array_equal(idx, want["indices"])


def test_pipeline_one_huge_segment_and_an_empty_class(backend):
    """A 40 000-pixel segment (38 752 colours: unique runs from the global workspace, stage 1 takes the
    MiniBatchKMeans branch, the split its wide-index form) next to a small one; the non-ROI class is empty."""
    H, W = 200, 260
    img = synth(H, W, 5, sigma=8.0)
    seg = np.ones((H, W), np.int32)
    seg[:, 200:] = 2
    roi = [{"bbox": (0, 0, H, W), "bbox_mask": np.ones((H, W), bool), "segments": seg}]
    want = O.encode_image(img, roi, [])
    pal, idx = _encode_device(backend, img, roi, [])
    assert np.array_equal(pal, want["palette"])
    assert np.array_equal(idx, want["indices"])


def test_host_encoder_streaming_equals_single_batches(backend):
    """HostEncoder.encode_many (upload of batch i+1, encode of batch i and download of batch i-1 overlap)
    returns, in order, exactly what encode returns batch by batch."""
    be = backend
    B, H, W = 2, 64, 96
    tab, lab = pipeline.table_from_tiles(B, H, W, 32)
    labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, B, H, W))))
    batches = [torch.from_numpy(np.stack([synth(H, W, 10 * s + i) for i in range(B)])) for s in range(4)]
    if be.device.type == "cuda":
        labs, batches = labs.pin_memory(), [b.pin_memory() for b in batches]
    enc = pipeline.HostEncoder(be, tab)
    outs = []
    for p, i in enc.encode_many([(b, labs) for b in batches]):
        outs.append((p, i))                                    # views: valid while the following batch is consumed
        if len(outs) >= 2:
            outs[-2] = (outs[-2][0], outs[-2][1].copy())
    outs = [(p, i.copy()) for p, i in outs]                    # encode() below reuses the result slots
    assert len(outs) == 4
    for s, (p, i) in enumerate(outs):
        p2, i2 = enc.encode(batches[s], labs)
        assert all(np.array_equal(a, b) for a, b in zip(p, p2)) and np.array_equal(i, i2), s
    roi, non = tile_regions(H, W, 32)
    want = O.encode_image(batches[3][1].numpy(), roi, non)
    assert np.array_equal(outs[3][0][1], want["palette"])
    assert np.array_equal(outs[3][1][1].reshape(-1).astype(np.int64), want["indices"])


def _oracle_frame(args):
    img, tile = args
    roi, non = tile_regions(img.shape[0], img.shape[1], tile)
    r = O.encode_image(img, roi, non)
    return r["palette"], r["indices"]


@pytest.mark.gpu
def test_full_hd_batch_equals_oracle_frame_by_frame():
    """BASELINE.json configs[1] at reduced batch: four 1920x1080 frames in one device batch, every frame
    compared with the oracle (run in parallel processes), palettes and index planes bit for bit."""
    import multiprocessing as mp
    from roibasedimagecompression_b200._lib import lib
    be = lib()
    B, H, W, tile = 4, 1080, 1920, 64
    imgs = np.stack([synth(H, W, 777 + i) for i in range(B)])
    tab, lab = pipeline.table_from_tiles(B, H, W, tile)
    labels = np.ascontiguousarray(np.broadcast_to(lab, (2, B, H, W)))
    res = pipeline.encode_batch(be, torch.from_numpy(imgs).cuda(), torch.from_numpy(labels).cuda(), tab)
    pipeline.finish_checks(res)
    with mp.get_context("fork").Pool(B) as pool:
        want = pool.map(_oracle_frame, [(imgs[b], tile) for b in range(B)])
    for b in range(B):
        assert np.array_equal(res.palette(b), want[b][0]), b
        assert np.array_equal(res.index_image(b).reshape(-1).astype(np.int64), want[b][1]), b
