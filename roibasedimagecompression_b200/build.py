"""In-tree nvcc build of librhccq.so for sm_100a (no JIT cache: the built file travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["rhccq_api.cu", "rhccq_palette.cu", "rhccq_split.cu", "rhccq_minibatch.cu", "rhccq_points.cu", "rhccq_pixels.cu", "rhccq_merge.cu", "rhccq_deflate.cu"]
OUT = os.path.join(HERE, "librhccq.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rhccq.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into librhccq.so (one nvcc per source, in parallel, then one link).
    Raises if nvcc is missing or a compile fails."""
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: librhccq.so must be built with the CUDA 12.9 toolkit for sm_100a")
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC"] + [o for o, _ in done] + ["-o", OUT + ".tmp"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(OUT + ".tmp", OUT)
    if verbose:
        print("".join(e for _, e in done))
    return OUT
