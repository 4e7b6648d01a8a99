#!/usr/bin/env python
"""Phase breakdown of rhccq_k_palette_split on a stage-1 batch: builds a profiling copy of the library with
-DRHCCQ_SPLIT_PROFILE (clock64 timers around the phases, summed over CTAs), runs one encode, prints shares.
Run on the GPU box:  python tools/split_phases.py [images]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from roibasedimagecompression_b200 import build as B, _lib, pipeline
from roibasedimagecompression_b200.synth import synth

out = os.path.join(ROOT, "gpurun_out", "librhccq_prof.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
nvcc = "/usr/local/cuda/bin/nvcc"
cmd = [nvcc] + B.NVCC_FLAGS + ["-DRHCCQ_SPLIT_PROFILE"] + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", out]
subprocess.run(cmd, check=True)
be = _lib.Backend(out, "cuda")
nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 1080, 1920
tab, lab = pipeline.table_from_tiles(nimg, H, W, 64)
imgs = torch.from_numpy(np.stack([synth(H, W, 1234 + i) for i in range(nimg)])).cuda()
labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, nimg, H, W)))).cuda()
pipeline.stage1(be, imgs, labs, tab)           # warm-up
torch.cuda.synchronize()
# zero / read the counters through a tiny helper exported by the profiling build
be.cdll.rhccq_split_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 8)()
be.cdll.rhccq_split_prof_read(buf, 1)           # reset
pipeline.stage1(be, imgs, labs, tab)
torch.cuda.synchronize()
be.cdll.rhccq_split_prof_read(buf, 0)
names = ["prologue", "top-level seeding (CTA)", "top-level Lloyd (CTA)", "CTA-level rest (partition, queue)",
         "warp-level splits (wall per CTA)", "leaf numbering"]
tot = sum(buf[i] for i in range(6)) or 1
v = [buf[i] for i in range(6)]
v[3] -= v[1] + v[2]                             # slot 3 spans the whole CTA-level phase
tot = sum(v)
for n, x in zip(names, v):
    print(f"{n:40s} {100.0 * x / tot:5.1f} %")
