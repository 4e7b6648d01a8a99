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

# the profiling library: rhccq_split.cu recompiled with the timers, linked with the product's other objects
# (build it in the build container with `python tools/split_phases.py --build-only`; it travels with the tree)
out = os.path.join(B.HERE, "build", "librhccq_prof.so")
if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(B.CSRC, "rhccq_split.cu")):
    B.build_library()
    nvcc = "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(B.HERE, "build")
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    subprocess.run([nvcc] + flags + ["-DRHCCQ_SPLIT_PROFILE", "-c", os.path.join(B.CSRC, "rhccq_split.cu"), "-o",
                    os.path.join(objdir, "rhccq_split_prof.o")], check=True)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in B.SOURCES if s != "rhccq_split.cu"]
    subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC"] + objs + [os.path.join(objdir, "rhccq_split_prof.o"), "-o", out], check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
be = _lib.Backend(out, "cuda")
nimg = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
H, W = 1080, 1920
tab, lab = pipeline.table_from_tiles(nimg, H, W, 64)
imgs = torch.from_numpy(np.stack([synth(H, W, 1234 + i) for i in range(nimg)])).cuda()
labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, nimg, H, W)))).cuda()
pipeline.stage1(be, imgs, labs, tab)           # warm-up
torch.cuda.synchronize()
# zero / read the counters through a tiny helper exported by the profiling build
be.cdll.rhccq_split_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
be.cdll.rhccq_split_prof_read(buf, 1)           # reset
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pipeline.stage1(be, imgs, labs, tab)
e1.record()
torch.cuda.synchronize()
be.cdll.rhccq_split_prof_read(buf, 0)
names = ["prologue", "top-level seeding (CTA)", "top-level Lloyd (CTA)", "CTA-level rest (partition, queue)",
         "warp-level splits (wall per CTA)", "leaf numbering"]
v = [buf[i] for i in range(6)]
v[3] -= v[1] + v[2]                             # slot 3 spans the whole CTA-level phase
tot = sum(v)
for n, x in zip(names, v):
    print(f"{n:40s} {100.0 * x / tot:5.1f} %")
print("E steps", buf[8], "; points through the bound test", buf[7], ", failing it", buf[9], ", still failing with the tightened bound", buf[10])
print("points to the second level", buf[14], "; E steps with third-level decisions", buf[15])
print("tolerance decisions in float64", buf[11], " seeding: draws re-evaluated", buf[12], " candidate ties", buf[13])
busy_ms = tot / 1.965e6
print(f"sum of the CTAs' phase times {busy_ms:.0f} ms; stage 1 took {e0.elapsed_time(e1):.1f} ms in all (split kernel ~85 % of it): "
      f"{busy_ms / 296:.1f} ms per resident CTA slot (296 = 2 per SM) -> what is missing to the kernel's time is the tail")
