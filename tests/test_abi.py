"""The C-ABI library: builds with nvcc for sm_100a, loads without a GPU, and exports every function that
include/rhccq.h declares; the ctypes table of the package mirrors the header one to one."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from roibasedimagecompression_b200 import _lib
from roibasedimagecompression_b200.build import build_library

HEADER = os.path.join(ROOT, "include", "rhccq.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rhccq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    path = build_library()                      # nvcc cross-compiles without a GPU; no-op when up to date
    cdll = ctypes.CDLL(path)
    missing = [n for n in _declared() if not hasattr(cdll, n)]
    assert not missing, missing
    cdll.rhccq_abi_version.restype = ctypes.c_int
    assert cdll.rhccq_abi_version() == 1        # host-only call: no GPU involved


def test_product_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.RhccqError):
        _lib.lib()
