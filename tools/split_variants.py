#!/usr/bin/env python
"""Time stage 1 (64 synthetic full-HD frames) with rhccq_split.cu compiled under different -D settings.
    python tools/split_variants.py --build-only tag1:-DX=1 tag2:-DX=2,-DY=3    (build container: nvcc)
    python tools/split_variants.py tag1 tag2                                      (GPU box)
`base` is the product library."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roibasedimagecompression_b200 import build as B
objdir = os.path.join(B.HERE, "build")


def lib_path(tag):
    return B.OUT if tag == "base" else os.path.join(objdir, f"librhccq_var_{tag}.so")


if "--build-only" in sys.argv:
    B.build_library()
    nvcc = "/usr/local/cuda/bin/nvcc"
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    procs = []
    for spec in [a for a in sys.argv[1:] if ":" in a]:
        tag, defs = spec.split(":", 1)
        obj = os.path.join(objdir, f"rhccq_split_var_{tag}.o")
        procs.append((tag, obj, subprocess.Popen([nvcc] + flags + defs.split(",") + ["-c", os.path.join(B.CSRC, "rhccq_split.cu"), "-o", obj])))
    for tag, obj, p in procs:
        assert p.wait() == 0, tag
        objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in B.SOURCES if s != "rhccq_split.cu"]
        subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC"] + objs + [obj, "-o", lib_path(tag)], check=True)
        print("built", lib_path(tag))
    sys.exit(0)

import numpy as np, torch
from roibasedimagecompression_b200 import _lib, pipeline
from roibasedimagecompression_b200.synth import synth
nimg, H, W = 64, 1080, 1920
tab, lab = pipeline.table_from_tiles(nimg, H, W, 64)
imgs = torch.from_numpy(np.stack([synth(H, W, 1234 + i) for i in range(nimg)])).cuda()
labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (2, nimg, H, W)))).cuda()
for tag in [a for a in sys.argv[1:] if not a.startswith("-")]:
    be = _lib.Backend(lib_path(tag), "cuda")
    pipeline.stage1(be, imgs, labs, tab)
    torch.cuda.synchronize()
    be.kernel_timing(True)
    for _ in range(3):
        st = pipeline.stage1(be, imgs, labs, tab)
    torch.cuda.synchronize()
    kt = be.kernel_times_ms()
    n, t = kt["rhccq_palette_split"]
    print(f"{tag:12s} rhccq_palette_split {t / n:8.2f} ms per launch; leaves checksum {int(st['nl1'].sum()) if hasattr(st['nl1'], 'sum') else 0}", flush=True)
