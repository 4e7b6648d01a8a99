"""The 'i' stream of the .rhccq package produced on the device (csrc/rhccq_deflate.cu).

The reference's writer compresses the index bytes with zlib level 9 on the host
(/root/reference/encoder/compression/compression.py:204-220); its reader only calls ``zlib.decompress``
(/root/reference/decoder/uncompression/uncompression.py:58-92).  For a batch whose index planes are already in
device memory, `zlib_streams` returns, per frame, a valid zlib stream of the plane's bytes built by the CUDA
library (fixed-Huffman blocks with run / row-above matches, one block per 4 KiB, Adler-32 combined here), and
`save_batch` wraps them into the reference's package (palette stream, pickle and the outer zlib on the host:
they are a few hundred KB per frame).  The files are read by the reference's reader and by ours; they are NOT
byte-identical to the reference's (only ``compression.save_encoded`` at level 9 is).
"""
from __future__ import annotations

import pickle
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import compression as _C

_ADLER = 65521


def zlib_streams(be, planes: torch.Tensor, elem_bytes: int, row_bytes: int) -> list:
    """planes: uint8 [F, n] on the device (the bytes of F index planes, ``elem_bytes`` per index, ``row_bytes``
    per image row).  Returns F ``bytes`` objects, each a zlib stream that decompresses to the plane's bytes."""
    if planes.dtype != torch.uint8 or planes.dim() != 2 or not planes.is_contiguous():
        raise ValueError("planes must be a contiguous uint8 [F, n] tensor")
    F, n = int(planes.shape[0]), int(planes.shape[1])
    if F == 0:
        return []
    ch, slot = int(be.cdll.rhccq_deflate_chunk_bytes()), int(be.cdll.rhccq_deflate_slot_bytes())
    cpf = (n + ch - 1) // ch
    head, tail = b"\x78\x01", b"\x03\x00"
    if cpf == 0:
        return [head + tail + struct.pack(">I", 1)] * F
    slots = be.empty((F * cpf * slot,), torch.uint8)
    slot_len = be.empty((F * cpf,), torch.int32)
    sums = be.empty((2 * F * cpf,), torch.int64)
    be.call("rhccq_deflate_chunks", be.ptr(planes), n, n, F, int(elem_bytes), int(row_bytes), be.ptr(slots),
            be.ptr(slot_len), be.ptr(sums), be.stream())
    lens = slot_len.cpu().numpy().astype(np.int64)
    s = sums.cpu().numpy().reshape(F, cpf, 2)
    offsets = np.zeros(F * cpf + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    out = be.empty((max(total, 1),), torch.uint8)
    d_off = torch.from_numpy(offsets[:-1].copy()).to(be.device)
    be.call("rhccq_deflate_pack", be.ptr(slots), be.ptr(slot_len), be.ptr(d_off), F * cpf, be.ptr(out), be.stream())
    body = out.cpu().numpy()
    # Adler-32 of a frame from the chunk sums: a = 1 + sum b_i, b = n + sum (n - i) b_i
    starts = (np.arange(cpf, dtype=np.int64) * ch)[None, :]
    a = (1 + s[:, :, 0].sum(axis=1)) % _ADLER
    b = (n + ((n - starts) * s[:, :, 0] - s[:, :, 1]).sum(axis=1)) % _ADLER
    res = []
    for f in range(F):
        lo, hi = int(offsets[f * cpf]), int(offsets[(f + 1) * cpf])
        res.append(head + body[lo:hi].tobytes() + tail + struct.pack(">I", (int(b[f]) << 16) | int(a[f])))
    return res


def package(palette_u8: np.ndarray, i_stream: bytes, shape, dtype_name: str, *, level: int = 1) -> bytes:
    """The bytes of one .rhccq file around a ready 'i' stream (compression.py:119-142, :151-202).  ``level`` is that
    of the outer zlib over the pickle.  Its content is the device's index stream, whose fixed Huffman code leaves
    redundancy: level 1 takes another 15 % off the file for ~0.2 ms per full-HD frame and thread (level 0 stores
    it: 0.99 instead of 0.84 bits per pixel on the bench frames).  The palette's few hundred bytes get level 9."""
    d = {"s": (int(shape[0]), int(shape[1])), "l": int(len(palette_u8)),
         "p": _C.compress_palette([tuple(int(v) for v in c) for c in palette_u8], 9), "i": i_stream, "d": dtype_name}
    body = zlib.compress(pickle.dumps(d, protocol=5), level)
    return _C.MAGIC + struct.pack("<I", len(body)) + body


def save_batch(be, result, filenames, *, level: int = 1, threads: int = 16) -> list:
    """Write every frame of an `EncodeResult` (pipeline.encode_batch) as a .rhccq file, the index streams built on
    the device.  Returns the file sizes.  Index dtype per frame as the reference chooses it (uint8 below 256
    palette rows, else uint16; compression.py:160-170)."""
    idx = result.indices
    B, H, W = (int(v) for v in idx.shape)
    if len(filenames) != B:
        raise ValueError("one file name per frame")
    cnt = result.palette_cnt.cpu().numpy()
    pals = [result.palette(b) for b in range(B)]
    streams, names = [None] * B, [None] * B
    narrow = [b for b in range(B) if cnt[b] <= 256]
    wide = [b for b in range(B) if cnt[b] > 256]
    if narrow:
        sel = idx if len(narrow) == B else idx[torch.as_tensor(narrow, device=idx.device)]
        planes = sel.to(torch.uint8).reshape(len(narrow), H * W).contiguous()
        for b, st in zip(narrow, zlib_streams(be, planes, 1, W)):
            streams[b], names[b] = st, "uint8"
    if wide:
        sel = idx if len(wide) == B else idx[torch.as_tensor(wide, device=idx.device)]
        planes = sel.contiguous().view(torch.uint8).reshape(len(wide), H * W * 2)
        for b, st in zip(wide, zlib_streams(be, planes, 2, 2 * W)):
            streams[b], names[b] = st, "uint16"

    def one(b):
        data = package(pals[b], streams[b], (H, W), names[b], level=level)
        with open(filenames[b], "wb") as f:
            f.write(data)
        return len(data)
    with ThreadPoolExecutor(max_workers=max(1, min(threads, B))) as ex:
        return list(ex.map(one, range(B)))
