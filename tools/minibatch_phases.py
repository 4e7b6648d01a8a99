#!/usr/bin/env python
"""Phase breakdown of rhccq_k_palette_minibatch on the 3840x2160 frame (BASELINE configs[2]): a profiling copy of
the library with -DRHCCQ_MB_PROFILE (clock64 timers of rank 0 / thread 0 of every cluster), one encode, shares.
Build it in the build container (`--build-only`; it travels with the tree), run on the GPU box."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roibasedimagecompression_b200 import build as B
out = os.path.join(B.HERE, "build", "librhccq_mbprof.so")
if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(B.CSRC, "rhccq_minibatch.cu")):
    B.build_library()
    nvcc = "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(B.HERE, "build")
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    subprocess.run([nvcc] + flags + ["-DRHCCQ_MB_PROFILE"] + [a for a in sys.argv[1:] if a.startswith("-D")] + ["-c", os.path.join(B.CSRC, "rhccq_minibatch.cu"), "-o",
                    os.path.join(objdir, "rhccq_minibatch_prof.o")], check=True)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in B.SOURCES if s != "rhccq_minibatch.cu"]
    subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC"] + objs + [os.path.join(objdir, "rhccq_minibatch_prof.o"), "-o", out], check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
import numpy as np, torch
from roibasedimagecompression_b200 import _lib, pipeline
from roibasedimagecompression_b200.synth import synth
be = _lib.Backend(out, "cuda")
H, W = 2160, 3840
tab, lab = pipeline.table_from_tiles(1, H, W, 64)
imgs = torch.from_numpy(synth(H, W, 1234)[None]).cuda()
labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, 1, H, W)))).cuda()
pipeline.encode_batch(be, imgs, labs, tab); torch.cuda.synchronize()
be.cdll.rhccq_mb_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 40)()
be.cdll.rhccq_mb_prof_read(buf, 1)
pipeline.encode_batch(be, imgs, labs, tab); torch.cuda.synchronize()
be.cdll.rhccq_mb_prof_read(buf, 0)
names = ["setup (rows, random subset)", "seeding", "cumulative probabilities", "batch rows (rank 0)", "  wait 1", "batch labels",
         "  wait 2", "centre update", "  wait 3", "inertia / reassignment / convergence (rank 0)", "  wait 4"]
tot = sum(buf[i] for i in range(11)) or 1
for i, nm in enumerate(names):
    print(f"{nm:50s} {100.0 * buf[i] / tot:5.1f} %   {buf[i] / 1.965e6:8.2f} ms summed over the clusters")
print("steps", buf[11], "of which with a reassignment", buf[14], "; sum of n", buf[12], "sum of k", buf[13])
print(f"longest palette: {buf[15] / 1.965e6:.2f} ms (the kernel lasts as long as its longest palette)")
print(f"batch points that needed the float64 level: {buf[16]} of {buf[11] * 1000} ; float64 scores evaluated for them: {buf[17]}")
print("batch labels per rank of the cluster, ms summed:", [round(buf[18 + r] / 1.965e6, 2) for r in range(8)])
