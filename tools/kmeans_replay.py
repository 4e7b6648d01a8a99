"""Replay every scikit-learn K-Means / MiniBatchKMeans call recorded from the reference.

``tests/golden/make_golden.py`` ran the reference's own modules and recorded, for every
``KMeans.fit_predict`` / ``MiniBatchKMeans.fit_predict`` they performed, the labels scikit-learn
returned (keys ``km_<sha1>`` / ``mb_<sha1>`` of the .npz fixtures; the hash covers colours and k).
This tool walks the same inputs through the oracle with those labels injected (which reproduces the
reference's outputs bit for bit), collects the (colours, k, labels) of every call on the way, and
compares each one with

* ``oracle/kmeans_restated.py`` (scikit-learn's arithmetic restated in C) and
  ``oracle/minibatch_restated.py``;
* optionally (``--sklearn``) scikit-learn itself, re-run here;
* optionally (``--device``) the CUDA kernel through the C ABI (needs a GPU).

    python tools/kmeans_replay.py [--sklearn] [--device] [--write]

``--write`` stores the summary as tests/golden/kmeans_replay.json (the table DESIGN.md quotes).
Test infrastructure: imports ``oracle``; nothing in the product imports this file.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FIXTURES = ["cluster_palette.npz", "pipeline_small.npz", "pipeline_synth.npz", "pipeline_lenna.npz",
            "minibatch_palette.npz"]


def lenna():
    from PIL import Image
    return np.array(Image.open(os.path.join(ROOT, "tests", "golden", "Lenna.png")).convert("RGB"))


def collect(fixtures=FIXTURES):
    """[(fixture, kind, colours uint8 [n,3], k, recorded labels)] in call order; every fixture's final
    output is asserted equal to the reference's on the way."""
    from conftest import golden, injected_kmeans, km_key
    from oracle import rhccq_oracle as O
    from roibasedimagecompression_b200.synth import tile_regions
    calls = []

    def recorder(g, tag):
        km = injected_kmeans(g)

        def f(colors, k):
            lab = km(colors, k)
            calls.append((tag, "km", np.asarray(colors).astype(np.uint8).copy(), int(k), lab.copy()))
            return lab
        return f

    for name in fixtures:
        g = golden(name)
        if name == "cluster_palette.npz":
            km = recorder(g, name)
            for c in range(int(g["n_cases"])):
                q = int(g[f"q{c}"])
                comp = {"palette": g[f"in_palette{c}"], "indices": g[f"in_indices{c}"],
                        "shape": tuple(g[f"shape{c}"]), "top_left": (0, 0)}
                eps, _, m = O.compute_clustering_params(len(comp["palette"]), q, "lab")
                r = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                      kmeans_impl=km)
                assert np.array_equal(r["palette"], g[f"out_palette{c}"]), (name, c)
                assert np.array_equal(r["indices"], g[f"out_indices{c}"]), (name, c)
        elif name == "minibatch_palette.npz":
            km = recorder(g, name)
            pal = g["in_palette"]
            table = {k[3:]: g[k] for k in g.files if k.startswith("mb_")}

            def mb(colors, k):
                lab = np.asarray(table[km_key(colors, k)]).astype(np.int64)
                calls.append((name, "mb", np.asarray(colors).astype(np.uint8).copy(), int(k), lab.copy()))
                return lab
            for c in range(int(g["n_cases"])):
                q = int(g[f"q{c}"])
                comp = {"palette": pal, "indices": np.arange(len(pal)), "shape": (1, len(pal)), "top_left": (0, 0)}
                eps, _, m = O.compute_clustering_params(len(pal), q, "lab")
                r = O.cluster_palette_colors_parallel(q, comp, eps=eps, min_samples=1, max_colors_per_cluster=m,
                                                      kmeans_impl=km, minibatch_impl=mb)
                assert np.array_equal(r["palette"], g[f"out_palette{c}"]), (name, c)
                assert np.array_equal(r["indices"], g[f"out_indices{c}"]), (name, c)
        else:
            img = g["image"] if "image" in g.files else lenna()
            roi, non = tile_regions(img.shape[0], img.shape[1], int(g["tile"]))
            r = O.encode_image(img, roi, non, kmeans_impl=recorder(g, name))
            assert np.array_equal(r["palette"], g["palette"]), name
            assert np.array_equal(r["indices"], g["indices"].astype(np.int64)), name
    return calls


def compare(calls, impl_km, impl_mb=None):
    """Per fixture: calls, exact calls, worst label agreement, and the diverging calls."""
    out = {}
    for tag, kind, col, k, lab in calls:
        impl = impl_km if kind == "km" else impl_mb
        if impl is None:
            continue
        s = out.setdefault(f"{tag}:{kind}", {"calls": 0, "exact": 0, "worst_agreement": 1.0, "diverging": []})
        got = np.asarray(impl(col, k)).astype(np.int64)
        s["calls"] += 1
        if np.array_equal(got, lab):
            s["exact"] += 1
        else:
            a = float((got == lab).mean())
            s["worst_agreement"] = min(s["worst_agreement"], a)
            s["diverging"].append({"n": int(len(col)), "k": int(k), "agreement": round(a, 6)})
    return out


def device_kmeans():
    """K-Means through the C ABI: one cluster holding every colour, split once (leaf == label)."""
    from roibasedimagecompression_b200._lib import lib
    from roibasedimagecompression_b200 import ops
    be = lib()

    def km(colors, k):
        return ops.kmeans_labels(be, colors, k)
    return km


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sklearn", action="store_true")
    ap.add_argument("--device", action="store_true")
    ap.add_argument("--write", action="store_true")
    a = ap.parse_args()
    from oracle import kmeans_restated as K, minibatch_restated as MB
    calls = collect()
    report = {"restated": compare(calls, K.kmeans_labels, MB.minibatch_labels)}
    if a.sklearn:
        import warnings
        from sklearn.cluster import KMeans, MiniBatchKMeans
        import sklearn

        def sk(col, k):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return KMeans(n_clusters=k, random_state=42, n_init="auto").fit_predict(col.astype(float))

        def skmb(col, k):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return MiniBatchKMeans(n_clusters=k, batch_size=1000, random_state=42,
                                       n_init="auto").fit_predict(col.astype(float))
        report["sklearn_rerun"] = compare(calls, sk, skmb)
        report["sklearn_version"] = sklearn.__version__
    if a.device:
        report["device"] = compare(calls, device_kmeans())
    tot = {k: {"calls": sum(v["calls"] for v in r.values()), "exact": sum(v["exact"] for v in r.values())}
           for k, r in report.items() if isinstance(r, dict)}
    report["totals"] = tot
    print(json.dumps(report, indent=1))
    if a.write:
        with open(os.path.join(ROOT, "tests", "golden", "kmeans_replay.json"), "w") as f:
            json.dump(report, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
