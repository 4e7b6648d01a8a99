#!/usr/bin/env python
"""Labels of scikit-learn's MiniBatchKMeans, as the reference calls it (clustering.py:207-218), on palettes of
10 000 colours and more -> tests/golden/minibatch_sklearn.npz.  Run in the build container (scikit-learn 1.9.0);
only palettes whose run does not pass through the unstable-argsort branch are kept (oracle/minibatch_restated.py)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sklearn.cluster import MiniBatchKMeans
from oracle import minibatch_restated as MB

rng = np.random.default_rng(2026)
out, kept = {}, 0
specs = [(14000, 12.0, 9), (11000, 30.0, 25), (16000, 6.0, 40), (12500, 55.0, 18), (10000, 80.0, 60)]
for n, sigma, q in specs:
    pts = np.clip(rng.integers(40, 215, 3) + rng.normal(size=(n * 3, 3)) * sigma, 1, 255).astype(np.uint8)
    col = np.unique(pts, axis=0)
    col = col[np.sort(rng.choice(len(col), min(n, len(col)), replace=False))]
    k = MB.n_clusters_for(len(col), q)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MiniBatchKMeans(n_clusters=k, batch_size=1000, random_state=42, n_init="auto").fit(col.astype(float))
    _, info = MB.minibatch_labels(col, k, return_info=True)
    print(len(col), q, k, "steps", m.n_steps_, "unstable cuts", info["unstable_cuts"])
    if info["unstable_cuts"]:
        continue
    out[f"colors{kept}"] = col
    out[f"q{kept}"] = np.float64(q)
    out[f"k{kept}"] = np.int64(k)
    out[f"labels{kept}"] = m.labels_.astype(np.int16)
    kept += 1
out["n_cases"] = np.int64(kept)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "minibatch_sklearn.npz"), **out)
print("kept", kept)
