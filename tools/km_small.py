#!/usr/bin/env python
"""A few K-Means calls through the C ABI (CTA-level and warp-level problems): the target of
`compute-sanitizer --tool racecheck` / `memcheck` runs.  python tools/km_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roibasedimagecompression_b200 import _lib, ops
from roibasedimagecompression_b200.encoder import compression as C
be = _lib.lib()
rng = np.random.default_rng(3)
for n, k in [(3000, 25), (700, 6), (200, 2), (1500, 140), (40, 40)]:
    col = np.unique(np.clip(rng.integers(40, 200, 3) + rng.normal(size=(n, 3)) * 14, 1, 255).astype(np.uint8), axis=0)
    lab = ops.kmeans_labels(be, col, min(k, len(col)))
    print(n, k, len(col), int(lab.max()) + 1)
pal = np.unique(np.clip(128 + rng.normal(size=(3500, 3)) * 25, 1, 255).astype(np.uint8), axis=0)
comp = {"palette": pal, "indices": np.arange(len(pal)), "shape": (1, len(pal)), "top_left": (0, 0)}
r = C.cluster_palette_colors_parallel(20, comp, eps=102.4, min_samples=1, max_colors_per_cluster=110, as_arrays=True)
print("split", len(pal), "->", len(r["palette"]))
