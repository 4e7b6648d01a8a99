#!/usr/bin/env python
"""Random palettes through the CUDA K-Means (rhccq_palette_split, max_cpc = -k) against the C restatement of
scikit-learn (oracle/kmeans_sklearn.c): the generator of tests/test_kernels_parity.py::_kmeans_cases with other
seeds and more cases.  Run on the GPU box:  python tools/kmeans_fuzz.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import kmeans_restated as K
from roibasedimagecompression_b200 import ops
from roibasedimagecompression_b200._lib import lib
from test_kernels_parity import _kmeans_cases, _ranks

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261019
be = lib()
rng = np.random.default_rng(seed)
t0 = time.time()
bad = done = reloc = 0
for col, k in _kmeans_cases(rng, cases):
    want, info = K.kmeans_labels(col, k, return_info=True)
    got = ops.kmeans_labels(be, col, k)
    done += 1
    reloc += info["relocations"] > 0
    if not np.array_equal(got, _ranks(want)):
        bad += 1
        print("MISMATCH", len(col), k, info, float((got == _ranks(want)).mean()), flush=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"kmeans_fuzz_bad_{bad}.npy"), np.concatenate([col.reshape(-1), [k]]))
print(f"{done} cases (seed {seed}), {reloc} with relocations, {bad} mismatches, {time.time() - t0:.0f} s")
