"""Generate the golden vectors under tests/golden/ from the reference itself.

Run ONCE in the build container, where the reference is mounted read-only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's own hot-path modules from /root/reference
(encoder.compression.{clustering,merging,regions,image}) unmodified, runs them
on small seeded inputs and stores inputs + outputs as compressed .npz files.
The GPU box has no /root/reference, so the tests read only these files.

Two things are normalised, both documented in SURVEY.md section 4 / 7.3:

* ``as_completed`` is replaced by submission order inside the reference's
  clustering module, removing its one nondeterministic step
  (clustering.py:458);
* every ``KMeans.fit_predict`` the reference performs is recorded
  (colours, k -> labels) so that a test can inject scikit-learn's own
  assignment into the implementation under test and compare everything
  downstream of K-Means bit for bit.

The stage-1 driver (encoder/compression/subregions.py) cannot be imported here
(it needs scikit-image); ``stage1_with_reference_ops`` below walks its
per-segment loop (:315-449, :634-683) with the label map supplied, calling the
reference's own get_all_unique_colors / compute_clustering_params /
cluster_palette_colors_parallel / merge_region_components_simple.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys
import threading
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

with contextlib.redirect_stdout(io.StringIO()):
    from encoder.compression import clustering as ref_clustering
    from encoder.compression import merging as ref_merging
    from encoder.compression import regions as ref_regions
    from encoder.compression import image as ref_image
import sklearn.cluster
from sklearn.cluster import DBSCAN
from PIL import Image

from roibasedimagecompression_b200.synth import synth, tile_regions   # deterministic inputs

ref_clustering.as_completed = lambda futures: list(futures)           # submission order

_KM_LOG: dict[str, np.ndarray] = {}
_KM_LOCK = threading.Lock()
_orig_fit_predict = sklearn.cluster.KMeans.fit_predict


def km_key(colors: np.ndarray, k: int) -> str:
    c = np.ascontiguousarray(np.asarray(colors).astype(np.uint8))
    return hashlib.sha1(c.tobytes() + int(k).to_bytes(4, "little")).hexdigest()


def _logged_fit_predict(self, X, y=None, sample_weight=None):
    labels = _orig_fit_predict(self, X, y, sample_weight)
    with _KM_LOCK:
        _KM_LOG[km_key(X, self.n_clusters)] = np.asarray(labels).astype(np.int32)
    return labels


sklearn.cluster.KMeans.fit_predict = _logged_fit_predict


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def arr_pal(comp):
    return np.asarray(comp["palette"], dtype=np.uint8).reshape(-1, 3)


def arr_idx(comp):
    return np.asarray(comp["indices"], dtype=np.int64).ravel()


def take_km_log() -> dict[str, np.ndarray]:
    with _KM_LOCK:
        out = dict(_KM_LOG)
        _KM_LOG.clear()
    return out


def stage1_with_reference_ops(image_rgb, regions, quality):
    """subregions.py:98-683 with ``region['segments']`` in place of SLIC."""
    out = []
    for region in regions:
        minr, minc, maxr, maxc = region["bbox"]
        region_image = image_rgb[minr:maxr, minc:maxc]
        bbox_mask = region["bbox_mask"]
        segs = region["segments"]
        comps = []
        ids = np.unique(segs)
        for sid in ids[ids != 0]:
            segment_mask = (segs == sid) & bbox_mask
            rows, cols = np.where(segment_mask)
            if len(rows) == 0:
                continue
            pad = 2
            h, w = region_image.shape[:2]
            r0, r1 = max(0, rows.min() - pad), min(h - 1, rows.max() + pad)
            c0, c1 = max(0, cols.min() - pad), min(w - 1, cols.max() + pad)
            crop = region_image[r0:r1 + 1, c0:c1 + 1]
            cm = segment_mask[r0:r1 + 1, c0:c1 + 1]
            seg_img = np.zeros_like(crop)
            seg_img[cm] = crop[cm]
            px = crop[cm]
            if len(px) and np.any(np.all(px == [0, 0, 0], axis=1)):
                nbm = ~np.all(px == [0, 0, 0], axis=1)
                nb = px[nbm]
                if len(nb):
                    for i in np.where(~nbm)[0]:
                        d = np.linalg.norm(nb - px[i], axis=1)
                        px[i] = nb[np.argmin(d)]
                seg_img[cm] = px
            comp = ref_clustering.get_all_unique_colors(seg_img, (int(r0 + minr), int(c0 + minc)))
            eps, _, mcpc = ref_clustering.compute_clustering_params(comp["actual_colors"], quality, "lab")
            comp = ref_clustering.cluster_palette_colors_parallel(
                quality, comp, eps=eps, min_samples=1, max_colors_per_cluster=mcpc)
            comps.append(comp)
        if len(comps) > 1:
            out.append(ref_merging.merge_region_components_simple(comps, (minr, minc, maxr, maxc)))
        else:
            out.append(comps)
    return out


def run_reference_pipeline(image_rgb, roi_regions, nonroi_regions, q_roi=20, q_non=10):
    """encoder/compression/test.py:100-142 on the reference's own functions."""
    H, W, _ = image_rgb.shape
    with quiet():
        s1r = stage1_with_reference_ops(image_rgb, roi_regions, q_roi)
        s1n = stage1_with_reference_ops(image_rgb, nonroi_regions, q_non)
        q2r, q2n = min(100, 2 * q_roi), min(100, 2 * q_non)
        roi = ref_regions.region_quantization(s1r, H, W, quality=q2r)
        non = ref_regions.region_quantization(s1n, H, W, quality=q2n)
        final = ref_image.quantize_image(roi + non, H, W, quality=min(100, q2r + q2n))
    stage_sizes = np.array([len(arr_pal(s1r[0][0])) if s1r and s1r[0] else 0,
                            len(arr_pal(s1n[0][0])) if s1n and s1n[0] else 0,
                            len(arr_pal(roi[0])), len(arr_pal(non[0])), len(arr_pal(final))])
    return final, stage_sizes, (s1r, s1n, roi, non)


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    lenna = np.array(Image.open(os.path.join(REF, "images/png/Lenna.png")).convert("RGB"))

    # ---- G1: DBSCAN labels at the reference's call site (clustering.py:205,233-235)
    d = {}
    tiles = [(0, 0), (3, 5), (6, 2)]
    qs = [20, 50, 60, 75, 90, 95, 97, 100]
    for ti, (ty, tx) in enumerate(tiles):
        t = lenna[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64].reshape(-1, 3)
        pal = np.unique(t, axis=0)
        d[f"pal{ti}"] = pal
        for q in qs:
            eps, ms, _ = ref_clustering.compute_clustering_params(len(pal), q)
            lab = DBSCAN(eps=eps / 255.0, min_samples=ms, metric="euclidean").fit_predict(
                pal.astype(float) / 255.0)
            d[f"lab{ti}_q{q}"] = lab.astype(np.int32)
    # a first-appearance-ordered (unsorted) palette, as stages 2/3 see it
    rng = np.random.default_rng(7)
    pal = np.unique(lenna[100:164, 200:264].reshape(-1, 3), axis=0)
    pal = pal[rng.permutation(len(pal))]
    d["pal3"] = pal
    for q in qs:
        eps, ms, _ = ref_clustering.compute_clustering_params(len(pal), q)
        d[f"lab3_q{q}"] = DBSCAN(eps=eps / 255.0, min_samples=ms).fit_predict(
            pal.astype(float) / 255.0).astype(np.int32)
    d["qs"] = np.array(qs)
    save("dbscan_palette.npz", **d)

    # ---- G1b: general DBSCAN (min_samples > 1) on 5-D pixel features, sklearn itself
    d = {}
    img = synth(48, 64, 1234)
    yy, xx = np.mgrid[0:48, 0:64]
    pts = np.concatenate([xx[..., None], yy[..., None], img], axis=2).reshape(-1, 5).astype(np.float32)
    d["image"] = img
    d["points"] = pts
    combos = [(3.0, 4), (5.0, 8), (8.0, 16), (2.0, 1), (4.5, 3)]
    for i, (eps, mp) in enumerate(combos):
        m = DBSCAN(eps=eps, min_samples=mp).fit(pts)
        d[f"labels{i}"] = m.labels_.astype(np.int32)
        core = np.zeros(len(pts), dtype=bool)
        core[m.core_sample_indices_] = True
        d[f"core{i}"] = core
    d["combos"] = np.array(combos, dtype=np.float64)
    u = np.random.default_rng(0).uniform(0, 24, size=(6000, 5)).astype(np.float32)
    d["uniform_points"] = u
    for i, (eps, mp) in enumerate([(3.0, 3), (4.0, 6)]):
        m = DBSCAN(eps=eps, min_samples=mp).fit(u)
        d[f"uniform_labels{i}"] = m.labels_.astype(np.int32)
    d["uniform_combos"] = np.array([(3.0, 3), (4.0, 6)], dtype=np.float64)
    save("dbscan_points.npz", **d)

    # ---- G2: get_all_unique_colors (clustering.py:4-103)
    d = {}
    crops = []
    c0 = lenna[10:40, 20:70].copy()
    c0[:5] = 0
    c0[:, :3] = 0
    crops.append(c0)
    crops.append(lenna[200:233, 300:331].copy())
    c2 = synth(24, 40, 5)
    c2[(np.add.outer(np.arange(24), np.arange(40)) % 5) == 0] = 0
    crops.append(c2)
    for i, c in enumerate(crops):
        with quiet():
            r = ref_clustering.get_all_unique_colors(c, (7 * i, 3 * i))
        d[f"crop{i}"] = c
        d[f"palette{i}"] = arr_pal(r)
        d[f"indices{i}"] = arr_idx(r).astype(np.int32)
    save("unique_colors.npz", **d)

    # ---- G3: cluster_palette_colors_parallel (clustering.py:160-437)
    d = {}
    take_km_log()
    case = 0
    for (ty, tx), with_black in [((1, 1), True), ((4, 6), False)]:
        t = lenna[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64].copy()
        if with_black:
            t[:2] = 0
            t[:, -2:] = 0
        with quiet():
            base = ref_clustering.get_all_unique_colors(t, (ty * 64, tx * 64))
        for q in [10, 20, 40, 60, 90, 95, 100]:
            eps, ms, mcpc = ref_clustering.compute_clustering_params(base["actual_colors"], q, "lab")
            with quiet():
                r = ref_clustering.cluster_palette_colors_parallel(
                    q, base, eps=eps, min_samples=1, max_colors_per_cluster=mcpc)
            d[f"in_palette{case}"] = arr_pal(base)
            d[f"in_indices{case}"] = arr_idx(base).astype(np.int32)
            d[f"shape{case}"] = np.array(base["shape"])
            d[f"q{case}"] = np.array(q)
            d[f"out_palette{case}"] = arr_pal(r)
            d[f"out_indices{case}"] = arr_idx(r).astype(np.int32)
            case += 1
    d["n_cases"] = np.array(case)
    for k, v in take_km_log().items():
        d["km_" + k] = v
    save("cluster_palette.npz", **d)

    # ---- G4: merge_region_components_simple (merging.py:8-120)
    d = {}
    rng = np.random.default_rng(11)
    comps = []
    boxes = [(2, 3, 10, 12), (6, 8, 9, 9), (0, 0, 5, 20), (12, 1, 6, 6), (-2, 15, 8, 8)]
    for (r, c, h, w) in boxes:
        m = int(rng.integers(2, 6))
        pal = rng.integers(0, 256, size=(m, 3)).astype(np.uint8)
        pal[0] = 0
        if m > 3:
            pal[3] = [9, 9, 9]          # a colour shared between components
        idx = rng.integers(0, m, size=h * w)
        comps.append({"top_left": (r, c), "shape": (h, w), "palette": pal.tolist(), "indices": idx.tolist()})
    bbox = (0, 0, 18, 22)
    with quiet():
        r = ref_merging.merge_region_components_simple(comps, bbox)[0]
    d["n_comp"] = np.array(len(comps))
    d["bbox"] = np.array(bbox)
    for i, c in enumerate(comps):
        d[f"top_left{i}"] = np.array(c["top_left"])
        d[f"shape{i}"] = np.array(c["shape"])
        d[f"palette{i}"] = arr_pal(c)
        d[f"indices{i}"] = arr_idx(c).astype(np.int32)
    d["out_palette"] = arr_pal(r)
    d["out_indices"] = arr_idx(r).astype(np.int32)
    d["out_shape"] = np.array(r["shape"])
    save("merge_canvas.npz", **d)

    # ---- G5 / G6: the three-stage pipeline (test.py:100-142) with tile segments
    for name, img, tile in [("pipeline_small.npz", lenna[192:288, 192:320].copy(), 32),
                            ("pipeline_synth.npz", synth(128, 192, 1234), 64),
                            ("pipeline_lenna.npz", lenna, 64)]:
        take_km_log()
        roi, non = tile_regions(img.shape[0], img.shape[1], tile)
        final, sizes, _ = run_reference_pipeline(img, roi, non)
        d = {"tile": np.array(tile), "stage_palette_sizes": sizes,
             "palette": arr_pal(final), "indices": arr_idx(final).astype(np.uint16),
             "shape": np.array(final["shape"])}
        if name != "pipeline_lenna.npz":
            d["image"] = img
        rec = arr_pal(final)[arr_idx(final).reshape(final["shape"])]
        mse = np.mean((rec.astype(np.float64) - img.astype(np.float64)) ** 2)
        d["psnr"] = np.array(10 * np.log10(255.0 ** 2 / mse))
        for k, v in take_km_log().items():
            d["km_" + k] = v.astype(np.int16)
        save(name, **d)
        print(f"  {name}: palette {len(arr_pal(final))}, psnr {float(d['psnr']):.2f} dB, stage sizes {sizes}")


def main_minibatch():
    """G7: the >= 10 000-colour branch (clustering.py:207-218): MiniBatchKMeans labels recorded."""
    _orig_mb = sklearn.cluster.MiniBatchKMeans.fit_predict
    log = {}

    def _logged(self, X, y=None, sample_weight=None):
        labels = _orig_mb(self, X, y, sample_weight)
        log[km_key(X, self.n_clusters)] = np.asarray(labels).astype(np.int32)
        return labels
    sklearn.cluster.MiniBatchKMeans.fit_predict = _logged
    ref_clustering.MiniBatchKMeans = sklearn.cluster.MiniBatchKMeans
    d = {}
    img = synth(256, 256, 3, sigma=3.0)
    pal = np.unique(img.reshape(-1, 3), axis=0)
    pal = pal[np.sort(np.random.default_rng(5).choice(len(pal), 10500, replace=False))]
    pal = np.concatenate([np.zeros((1, 3), np.uint8), pal])
    take_km_log()
    for case, q in enumerate([10, 40]):
        comp = {"palette": pal.tolist(), "indices": list(range(len(pal))), "shape": (1, len(pal)), "top_left": (0, 0)}
        eps, ms, mcpc = ref_clustering.compute_clustering_params(len(pal), q, "lab")
        with quiet():
            r = ref_clustering.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=mcpc)
        d[f"q{case}"] = np.array(q)
        d[f"out_palette{case}"] = arr_pal(r)
        d[f"out_indices{case}"] = arr_idx(r).astype(np.int32)
    d["in_palette"] = pal
    d["n_cases"] = np.array(2)
    for k, v in take_km_log().items():
        d["km_" + k] = v.astype(np.int16)
    for k, v in log.items():
        d["mb_" + k] = v.astype(np.int16)
    save("minibatch_palette.npz", **d)


def main_container():
    """G8: the container writer (compression.py:119-220) on the final result of the small pipeline golden,
    and one of the reference's own shipped outputs (images/rhccq_20_10/Lenna_compressed.rhccq) for the reader."""
    import shutil
    import tempfile
    with quiet():
        from encoder.compression import compression as ref_comp
    g = np.load(os.path.join(OUT, "pipeline_small.npz"))
    pal, idx, shape = g["palette"], g["indices"], tuple(int(v) for v in g["shape"])
    with quiet():
        pkg = ref_comp.lossless_compress_optimized([tuple(int(v) for v in c) for c in pal],
                                                   idx.astype(np.int64).reshape(shape), shape)
        fn = os.path.join(tempfile.mkdtemp(), "x.rhccq")
        ref_comp.save_compressed(pkg, fn)
    shutil.copy(fn, os.path.join(OUT, "container_small.rhccq"))
    shutil.copy(os.path.join(REF, "images/rhccq_20_10/Lenna_compressed.rhccq"), os.path.join(OUT, "reference_Lenna_compressed.rhccq"))
    print("wrote container_small.rhccq", os.path.getsize(os.path.join(OUT, "container_small.rhccq")), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "container":
        main_container()
    elif len(sys.argv) > 1 and sys.argv[1] == "minibatch":
        main_minibatch()
    else:
        main()
