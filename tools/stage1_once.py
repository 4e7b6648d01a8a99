#!/usr/bin/env python
"""One stage-1 pass (unique, DBSCAN, split, finish, remap) over a few synthetic 1920x1080 frames: the
target of `ncu --kernel-name regex:rhccq_k_palette_split` captures (see profiles/).  python tools/stage1_once.py [images] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from roibasedimagecompression_b200 import _lib, pipeline
from roibasedimagecompression_b200.synth import synth
be = _lib.lib()
nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W = 1080, 1920
tab, lab = pipeline.table_from_tiles(nimg, H, W, 64)
imgs = torch.from_numpy(np.stack([synth(H, W, 1234 + i) for i in range(nimg)])).cuda()
labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, nimg, H, W)))).cuda()
for _ in range(reps):
    pipeline.stage1(be, imgs, labs, tab)
torch.cuda.synchronize()
print("ok")
