#!/usr/bin/env python
"""Random small images, tile sizes and quality triples through pipeline.encode_batch on the GPU against the oracle
(oracle/rhccq_oracle.py with the restated K-Means): palette and index plane compared exactly.
Run on the GPU box:  python tools/pipeline_fuzz.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import rhccq_oracle as O
from roibasedimagecompression_b200 import pipeline
from roibasedimagecompression_b200._lib import lib, RhccqError
from roibasedimagecompression_b200.synth import synth, tile_regions

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31337
be = lib()
rng = np.random.default_rng(seed)
bad = refused = 0
t0 = time.time()
for c in range(cases):
    H, W = int(rng.integers(3, 9)) * 16, int(rng.integers(3, 12)) * 16
    tile = int(rng.choice([16, 32, 48, 64]))
    sigma = float(rng.choice([0.0, 1.0, 3.0, 6.0]))
    img = synth(H, W, int(rng.integers(0, 10 ** 6)), sigma=sigma) if sigma else synth(H, W, int(rng.integers(0, 10 ** 6)))
    if rng.random() < 0.3:
        img = (img // int(rng.integers(2, 9))) * int(rng.integers(1, 5))        # flatter: fewer colours, more ties
        img = np.clip(img, 0, 255).astype(np.uint8)
    q1 = int(rng.choice([10, 20, 30, 50])); q0 = int(rng.choice([5, 10, 20]))
    quals = (q1, q0)
    roi, non = tile_regions(H, W, tile)
    try:
        want = O.encode_image(img, roi, non, roi_quality=quals[0], nonroi_quality=quals[1])
    except Exception as e:                                       # the oracle refuses what the reference cannot do either
        want = e
    tab, lab = pipeline.table_from_regions((H, W), [roi, non], quals)
    try:
        res = pipeline.encode_batch(be, torch.from_numpy(img[None].copy()).cuda(), torch.from_numpy(lab).cuda(), tab)
        pipeline.finish_checks(res)
        got = (res.palette(0), res.index_image(0).reshape(-1).astype(np.int64))
    except RhccqError as e:
        got = e
    if isinstance(want, Exception) or isinstance(got, Exception):
        refused += 1
        print(f"{H}x{W} tile {tile} q {quals}: oracle {type(want).__name__ if isinstance(want, Exception) else 'ok'}, "
              f"device {(type(got).__name__ + ': ' + str(got)[:160]) if isinstance(got, Exception) else 'ok'}", flush=True)
        continue
    ok = np.array_equal(got[0], want["palette"]) and np.array_equal(got[1], np.asarray(want["indices"]).reshape(-1))
    bad += not ok
    print(f"{H}x{W} tile {tile} sigma {sigma} q {quals}: {len(want['palette'])} colours, {'equal' if ok else 'MISMATCH'}", flush=True)
print(f"{bad} mismatches, {refused} refused, {time.time() - t0:.0f} s")
