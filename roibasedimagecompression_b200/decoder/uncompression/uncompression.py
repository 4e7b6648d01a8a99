"""Counterpart of /root/reference/decoder/uncompression/uncompression.py.

`load_compressed` / `lossless_decompress` read the reference's container (uncompression.py:129-150, :58-92);
`decompress_color_quantization` is the palette gather (`palette[indices]`, :156-218) and `quality_metrics`
the MSE / PSNR / MAE of decoder/uncompression/comparison.py:43-44,64-79 — both on the GPU
(rhccq_decode_gather, rhccq_sq_abs_err), so that an encode can be verified in the same run without a host
round trip of the pixels.

The reference unpickles with `pickle.loads` (uncompression.py:150), which executes arbitrary globals of an
untrusted file; this reader only admits the two numpy globals that the reference's own files contain.
"""
from __future__ import annotations

import io
import pickle
import struct
import zlib

import numpy as np
import torch

from ..._lib import lib, RhccqError

MAGIC = b"RHCCQ"
_BACKEND = None          # tests bind the host-emulation build here


def _be():
    return _BACKEND if _BACKEND is not None else lib()


class _Restricted(pickle.Unpickler):
    ALLOWED = {("numpy._core.multiarray", "scalar"), ("numpy.core.multiarray", "scalar"), ("numpy", "dtype")}

    def find_class(self, module, name):
        if (module, name) in self.ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"global {module}.{name} is not allowed in an .rhccq file")


def load_compressed(filename) -> dict:
    """uncompression.py:129-150."""
    with open(filename, "rb") as f:
        if f.read(5) != MAGIC:
            raise ValueError("Invalid file format")
        size = struct.unpack("<I", f.read(4))[0]
        body = f.read(size)
    return _Restricted(io.BytesIO(zlib.decompress(body))).load()


def lossless_decompress(compressed_data: dict):
    """uncompression.py:58-92 -> (palette uint8 [l,3], indices [h*w] in the stored dtype, shape)."""
    h, w = (int(v) for v in compressed_data["s"])
    n_pal = int(compressed_data["l"])
    dtype = np.dtype(compressed_data.get("d", "uint16"))            # default of the reference: uint16
    palette = np.frombuffer(zlib.decompress(compressed_data["p"]), dtype=np.uint8).reshape(n_pal, 3)
    indices = np.frombuffer(zlib.decompress(compressed_data["i"]), dtype=dtype)
    if indices.size != h * w:
        raise ValueError(f"{indices.size} indices for shape {(h, w)}")
    return palette, indices, (h, w)


def decompress_color_quantization(palette, indices, shape, *, device_result: bool = False):
    """uncompression.py:156-218: image[h,w,3] = palette[indices].  The gather runs on the GPU."""
    be = _be()
    h, w = shape
    pal = torch.from_numpy(np.ascontiguousarray(palette, dtype=np.uint8).reshape(-1, 3)).to(be.device)
    idx_np = np.ascontiguousarray(indices).reshape(-1)
    if idx_np.dtype not in (np.uint8, np.uint16, np.uint32):
        idx_np = idx_np.astype(np.uint32)
    idx = torch.from_numpy(idx_np.view({1: np.uint8, 2: np.int16, 4: np.int32}[idx_np.dtype.itemsize])).to(be.device)
    out = be.empty((h, w, 3), torch.uint8)
    bad = be.zeros((1,), torch.int32)
    be.call("rhccq_decode_gather", be.ptr(idx), int(idx_np.dtype.itemsize), int(h * w), be.ptr(pal), int(pal.shape[0]),
            be.ptr(out), be.ptr(bad), be.stream())
    if int(bad.item()):
        raise RhccqError("index outside the palette")
    return out if device_result else out.cpu().numpy()


def quality_metrics(a, b) -> dict:
    """MSE / PSNR / MAE of two uint8 images (comparison.py:43-44,64-79), reduced on the GPU."""
    be = _be()
    ta = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8))
    tb = b if isinstance(b, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(b, dtype=np.uint8))
    ta, tb = ta.to(be.device).contiguous(), tb.to(be.device).contiguous()
    if ta.shape != tb.shape or ta.dtype != torch.uint8 or tb.dtype != torch.uint8:
        raise ValueError("two uint8 images of the same shape are required")
    acc = be.zeros((2,), torch.int64)
    be.call("rhccq_sq_abs_err", be.ptr(ta), be.ptr(tb), int(ta.numel()), be.ptr(acc), be.stream())
    sq, ab = (int(v) for v in acc.cpu().tolist())
    n = max(int(ta.numel()), 1)
    mse = sq / n
    return {"mse": mse, "psnr": float("inf") if mse == 0 else float(10 * np.log10(255.0 * 255.0 / mse)), "mae": ab / n}
