import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_SO = os.path.join(ROOT, "tests", "emu", "_build", "librhccq_emu.so")
CSRC = os.path.join(ROOT, "roibasedimagecompression_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def km_key(colors, k):
    c = np.ascontiguousarray(np.asarray(colors).astype(np.uint8))
    return hashlib.sha1(c.tobytes() + int(k).to_bytes(4, "little")).hexdigest()


def injected_kmeans(g):
    """K-Means callable that replays scikit-learn's own labels recorded by make_golden.py."""
    table = {k[3:]: g[k] for k in g.files if k.startswith("km_")}

    def km(colors, k):
        return np.asarray(table[km_key(colors, k)]).astype(np.int64)
    return km


def _emu_stale():
    if not os.path.exists(EMU_SO):
        return True
    t = os.path.getmtime(EMU_SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in os.listdir(CSRC))


@pytest.fixture(scope="session")
def emu_backend():
    """Host-emulation build of the kernel sources (test infrastructure, never the product path)."""
    if _emu_stale():
        subprocess.run(["sh", os.path.join(ROOT, "tests", "emu", "build_emu.sh")], check=True,
                       capture_output=True)
    from roibasedimagecompression_b200._lib import Backend
    return Backend(EMU_SO, "cpu")


@pytest.fixture(scope="session")
def emu_exact_backend(emu_backend):
    """The emulation build with -DRHCCQ_KM_FORCE_EXACT: every K-Means decision takes the float64 path."""
    from roibasedimagecompression_b200._lib import Backend
    return Backend(EMU_SO.replace("librhccq_emu.so", "librhccq_emu_exact.so"), "cpu")


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """The kernels under test: the emulation build on CPU, the nvcc build on the GPU box."""
    from roibasedimagecompression_b200.encoder.compression import clustering
    if request.param == "emu":
        be = request.getfixturevalue("emu_backend")
    else:
        from roibasedimagecompression_b200._lib import lib
        be = lib()                       # raises when librhccq.so or the GPU is missing
    clustering._BACKEND = be if request.param == "emu" else None
    yield be
    clustering._BACKEND = None


def same_component(a, b):
    pa = np.asarray(a["palette"], dtype=np.int64).reshape(-1, 3)
    pb = np.asarray(b["palette"], dtype=np.int64).reshape(-1, 3)
    return (np.array_equal(pa, pb)
            and np.array_equal(np.asarray(a["indices"]).ravel(), np.asarray(b["indices"]).ravel())
            and tuple(a["shape"]) == tuple(b["shape"]) and tuple(a["top_left"]) == tuple(b["top_left"]))
