#!/usr/bin/env python
"""Random palettes of 10 000 colours and more through rhccq_palette_minibatch (ops.minibatch_labels) against
oracle/minibatch_restated.py, labels compared exactly; includes k > 1 000, where a call passes through the
stable-order reassignment cut.  Run on the GPU box:  python tools/minibatch_fuzz.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import minibatch_restated as MB
from roibasedimagecompression_b200 import ops
from roibasedimagecompression_b200._lib import lib

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
be = lib()
rng = np.random.default_rng(seed)
bad = 0
t0 = time.time()
for c in range(cases):
    n = int(rng.integers(10000, 30000))
    pts = np.clip(rng.integers(40, 215, 3) + rng.normal(size=(n * 3, 3)) * rng.uniform(8, 60), 1, 255).astype(np.uint8)
    col = np.unique(pts, axis=0)
    col = col[np.sort(rng.choice(len(col), min(n, len(col)), replace=False))]
    if len(col) < 10000:
        continue
    q = float(rng.integers(5, 100))
    k = MB.n_clusters_for(len(col), q)
    want, info = MB.minibatch_labels(col, k, return_info=True)
    got, k_dev = ops.minibatch_labels(be, col, q)
    ok = k_dev == k and np.array_equal(got, want)
    bad += not ok
    print(f"n {len(col)} q {q} k {k} steps {info['steps']} unstable cuts {info['unstable_cuts']}: {'equal' if ok else 'MISMATCH'}", flush=True)
print(f"{bad} mismatches, {time.time() - t0:.0f} s")
