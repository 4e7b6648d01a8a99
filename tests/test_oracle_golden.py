"""The oracle against the golden vectors produced by the reference itself (tests/golden/make_golden.py).

These pin oracle/: every function of the restatement must reproduce what the reference's own modules
(and scikit-learn, the third-party code they call) returned on the same inputs.
"""
import numpy as np
import pytest

from conftest import golden, injected_kmeans
from oracle import rhccq_oracle as O, kmeans_restated as K


def test_dbscan_labels_match_sklearn_at_reference_call_site():
    g = golden("dbscan_palette.npz")
    for ti in range(4):
        pal = g[f"pal{ti}"]
        for q in g["qs"]:
            eps, ms, _ = O.compute_clustering_params(len(pal), int(q))
            lab = O.dbscan_labels(pal, eps, ms, colour_scale=True)
            assert np.array_equal(lab, g[f"lab{ti}_q{q}"]), (ti, int(q))


def test_dbscan_points_general_minpts():
    g = golden("dbscan_points.npz")
    pts = g["points"]
    for i, (eps, mp) in enumerate(g["combos"]):
        lab = O.dbscan_labels(pts, float(eps), int(mp))
        assert np.array_equal(lab, g[f"labels{i}"]), (eps, mp)
    u = g["uniform_points"]
    for i, (eps, mp) in enumerate(g["uniform_combos"]):
        assert np.array_equal(O.dbscan_labels(u, float(eps), int(mp)), g[f"uniform_labels{i}"])


def test_unique_colors():
    g = golden("unique_colors.npz")
    for i in range(3):
        r = O.get_all_unique_colors(g[f"crop{i}"], (7 * i, 3 * i))
        assert np.array_equal(r["palette"], g[f"palette{i}"])
        assert np.array_equal(r["indices"], g[f"indices{i}"])


def test_cluster_palette_with_sklearn_kmeans_injected():
    g = golden("cluster_palette.npz")
    km = injected_kmeans(g)
    for c in range(int(g["n_cases"])):
        q = int(g[f"q{c}"])
        comp = {"palette": g[f"in_palette{c}"], "indices": g[f"in_indices{c}"],
                "shape": tuple(g[f"shape{c}"]), "top_left": (0, 0)}
        eps, _, m = O.compute_clustering_params(len(comp["palette"]), q, "lab")
        r = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                              kmeans_impl=km)
        assert np.array_equal(r["palette"], g[f"out_palette{c}"]), c
        assert np.array_equal(r["indices"], g[f"out_indices{c}"]), c


def test_merge_canvas():
    g = golden("merge_canvas.npz")
    comps = [{"top_left": tuple(g[f"top_left{i}"]), "shape": tuple(g[f"shape{i}"]),
              "palette": g[f"palette{i}"], "indices": g[f"indices{i}"]} for i in range(int(g["n_comp"]))]
    r = O.merge_region_components_simple(comps, tuple(g["bbox"]))[0]
    assert np.array_equal(r["palette"], g["out_palette"])
    assert np.array_equal(r["indices"], g["out_indices"])
    assert tuple(r["shape"]) == tuple(g["out_shape"])


@pytest.mark.parametrize("name", ["pipeline_small.npz", "pipeline_synth.npz"])
def test_pipeline_with_sklearn_kmeans_injected(name):
    from roibasedimagecompression_b200.synth import tile_regions
    g = golden(name)
    img = g["image"]
    roi, non = tile_regions(img.shape[0], img.shape[1], int(g["tile"]))
    r = O.encode_image(img, roi, non, kmeans_impl=injected_kmeans(g))
    assert np.array_equal(r["palette"], g["palette"])
    assert np.array_equal(r["indices"], g["indices"].astype(np.int64))


def _pipeline_image(g):
    if "image" in g.files:
        return g["image"]
    from PIL import Image
    import os
    from conftest import GOLDEN
    return np.array(Image.open(os.path.join(GOLDEN, "Lenna.png")).convert("RGB"))   # BASELINE configs[0]


@pytest.mark.parametrize("name", ["pipeline_small.npz", "pipeline_synth.npz", "pipeline_lenna.npz"])
def test_pipeline_restated_kmeans_equals_reference(name):
    """No injection: with scikit-learn's K-Means restated in its own arithmetic (oracle/kmeans_sklearn.c) the
    oracle reproduces the reference's palette and index plane bit for bit — a Lenna crop, a synthetic image
    and BASELINE configs[0] itself (Lenna 512x512 at rhccq_20_10, 729 K-Means calls)."""
    from roibasedimagecompression_b200.synth import tile_regions
    g = golden(name)
    img = _pipeline_image(g)
    roi, non = tile_regions(img.shape[0], img.shape[1], int(g["tile"]))
    r = O.encode_image(img, roi, non)
    assert np.array_equal(r["palette"], g["palette"])
    assert np.array_equal(r["indices"], g["indices"].astype(np.int64))
    assert abs(O.psnr(O.decode(r), img) - float(g["psnr"])) < 1e-9


def test_every_recorded_sklearn_kmeans_call_is_reproduced():
    """tools/kmeans_replay.py: the 3 693 KMeans and 2 MiniBatchKMeans calls scikit-learn performed while
    make_golden.py ran the reference, label for label."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import kmeans_replay
    from oracle import minibatch_restated as MB
    calls = kmeans_replay.collect()
    rep = kmeans_replay.compare(calls, K.kmeans_labels, MB.minibatch_labels)
    assert sum(v["calls"] for v in rep.values()) == 3695
    for tag, v in rep.items():
        assert v["exact"] == v["calls"], (tag, v["diverging"][:3])


def test_first_seed_is_floor_of_u_times_n():
    """RandomState(42).choice(n, p=uniform) (cumulative sum of 1/n, normalised, searchsorted right) equals
    floor(u n) for the first uniform of the stream and every n the kernel can see — the kernel uses the
    closed form."""
    u = float(K.rng_doubles(1)[0])
    for n in range(1, 20001):
        assert K.first_seed(n) == min(int(u * n), n - 1), n


def test_kmeans_restated_equals_sklearn_on_random_palettes():
    """Where scikit-learn is installed: labels identical on random palettes of every regime (tiny with k up
    to n, image-like, k > 192 where Lloyd's dgemm is blocked, duplicate colours with empty-cluster
    relocation)."""
    pytest.importorskip("sklearn")
    import warnings
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(123)
    cases = []
    for t in range(60):
        n = int(rng.integers(3, 400))
        col = np.unique(rng.integers(0, 256, (n, 3)).astype(np.uint8), axis=0)
        cases.append((col, int(rng.integers(2, min(len(col), 40) + 1))))
    for t in range(30):
        n = int(rng.integers(5, 3000))
        col = np.unique(np.clip(rng.integers(30, 220, 3) + rng.normal(size=(n, 3)) * rng.uniform(2, 25), 0, 255)
                        .astype(np.uint8), axis=0)
        cases.append((col, int(rng.integers(2, min(len(col), 150) + 1))))
    for t in range(80):
        col = np.unique(rng.integers(100, 110, (int(rng.integers(2, 30)), 3)).astype(np.uint8), axis=0)
        cases.append((col, int(rng.integers(1, len(col) + 1))))
    for t in range(6):
        col = np.unique(rng.integers(0, 64, (int(rng.integers(600, 2500)), 3)).astype(np.uint8), axis=0)
        cases.append((col, int(rng.integers(193, len(col) // 2))))
    reloc = 0
    for t in range(60):
        n = int(rng.integers(6, 80))
        base = rng.integers(1, 255, (int(rng.integers(2, 8)), 3))
        col = base[rng.integers(0, len(base), n)].astype(np.uint8)             # duplicates: clusters go empty
        cases.append((col, int(rng.integers(2, min(n, 12) + 1))))
    for col, k in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sk = KMeans(n_clusters=k, random_state=42, n_init="auto").fit_predict(col.astype(float))
        lab, info = K.kmeans_labels(col, k, return_info=True)
        reloc += info["relocations"] > 0
        assert np.array_equal(lab, sk), (len(col), k, info)
    assert reloc > 5


def test_minibatch_branch_with_sklearn_labels_injected():
    """clustering.py:207-218 through the reference itself (tests/golden/make_golden.py minibatch): with
    scikit-learn's recorded MiniBatchKMeans / KMeans labels injected the oracle reproduces the reference's
    palette (order included) and indices."""
    from conftest import km_key
    g = golden("minibatch_palette.npz")
    pal = g["in_palette"]
    km = injected_kmeans(g)
    mb_table = {k[3:]: g[k] for k in g.files if k.startswith("mb_")}

    def mb(colors, k):
        return np.asarray(mb_table[km_key(colors, k)]).astype(np.int64)
    for c in range(int(g["n_cases"])):
        q = int(g[f"q{c}"])
        comp = {"palette": pal, "indices": np.arange(len(pal)), "shape": (1, len(pal)), "top_left": (0, 0)}
        eps, _, m = O.compute_clustering_params(len(pal), q, "lab")
        r = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                              kmeans_impl=km, minibatch_impl=mb)
        assert np.array_equal(r["palette"], g[f"out_palette{c}"]), c
        assert np.array_equal(r["indices"], g[f"out_indices{c}"]), c


def test_minibatch_restated_equals_sklearn():
    """Where scikit-learn is installed: labels and final centres identical to MiniBatchKMeans itself — the golden
    palette at two qualities and random palettes — whenever the call does not pass through the one branch whose
    outcome in scikit-learn depends on numpy's unstable argsort (oracle/minibatch_restated.py)."""
    pytest.importorskip("sklearn")
    import warnings
    from sklearn.cluster import MiniBatchKMeans
    from oracle import minibatch_restated as MB
    g = golden("minibatch_palette.npz")
    pal = g["in_palette"][1:]
    cases = [(pal, MB.n_clusters_for(len(pal), q)) for q in (10, 40)]
    rng = np.random.default_rng(77)
    for t in range(3):
        n = int(rng.integers(10000, 30000))
        col = np.unique(np.clip(rng.integers(30, 220, 3) + rng.normal(size=(n, 3)) * rng.uniform(6, 40), 0, 255)
                        .astype(np.uint8), axis=0)
        cases.append((col, MB.n_clusters_for(len(col), int(rng.integers(5, 30)))))
    stable = 0
    for col, k in cases:
        lab, info = MB.minibatch_labels(col, k, return_info=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = MiniBatchKMeans(n_clusters=k, batch_size=1000, random_state=42, n_init="auto").fit(col.astype(float))
        if info["unstable_cuts"]:
            continue
        stable += 1
        assert info["steps"] == m.n_steps_
        assert np.array_equal(info["centers"], m.cluster_centers_), (len(col), k)
        assert np.array_equal(lab, m.labels_), (len(col), k)
    assert stable >= 4
