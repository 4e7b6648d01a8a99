"""Counterpart of /root/reference/encoder/compression/compression.py (container writer).

The byte format is the reference's and is kept as it is, so that its decoder
(decoder/uncompression/uncompression.py:129-150, :58-92) reads the files:

    b'RHCCQ' | uint32 little-endian length | zlib9( pickle5( {'s': (h, w), 'l': palette rows,
        'p': zlib9(palette uint8 bytes), 'i': zlib9(indices in the narrowest unsigned dtype), 'd': dtype name} ) )

(compression.py:10-22 palette, :151-220 package, :119-142 file).  zlib and pickle are the same library calls
the reference makes; nothing here is on the GPU path — `save_encoded` takes the final palette and index plane
of `pipeline.encode_batch` / `HostEncoder` (already on the host) and writes the file.
"""
from __future__ import annotations

import pickle
import struct
import zlib

import numpy as np

MAGIC = b"RHCCQ"


def compress_palette(palette, level: int = 9) -> bytes:
    """compression.py:10-22."""
    return zlib.compress(np.array(palette, dtype=np.uint8).tobytes(), level=level)


def compress_indices_simple_optimized(indices_data, dtype=np.uint8, level: int = 9) -> bytes:
    """compression.py:204-220."""
    idx = indices_data if isinstance(indices_data, np.ndarray) else np.array(indices_data, dtype=dtype)
    if idx.dtype != dtype:
        idx = idx.astype(dtype)
    return zlib.compress(idx.tobytes(), level=level)


def lossless_compress_optimized(palette, indices_list, shape, use_manual_rle=False, *, level: int = 9) -> dict:
    """compression.py:151-202: the narrowest of uint8 / uint16 / uint32 that holds max(indices).

    ``level`` is the zlib level of all three streams.  The reference writes level 9 (the default here: the
    file is then byte-identical to the reference's); its reader (decoder/uncompression/uncompression.py:58-92)
    only calls ``zlib.decompress`` and accepts any level, so a batch encoder whose host cores are the
    bottleneck can trade a few per cent of size for most of the time."""
    if isinstance(indices_list, np.ndarray):
        max_index = int(indices_list.max()) if indices_list.size else 0
        flat = indices_list.flatten()
    elif isinstance(indices_list, list):
        max_index = max(indices_list) if indices_list else 0
        flat = indices_list
    else:
        raise TypeError(f"indices_list must be list or numpy array, got {type(indices_list)}")
    if max_index < 256:
        dtype, name = np.uint8, "uint8"
    elif max_index < 65536:
        dtype, name = np.uint16, "uint16"
    else:
        dtype, name = np.uint32, "uint32"
    return {"s": shape, "l": len(palette), "p": compress_palette(palette, level),
            "i": compress_indices_simple_optimized(flat, dtype, level), "d": name}


def save_compressed(compressed_data: dict, filename, *, level: int = 9) -> int:
    """compression.py:119-142.  Returns the file size in bytes."""
    body = zlib.compress(pickle.dumps(compressed_data, protocol=5), level=level)
    with open(filename, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(body)))
        f.write(body)
    return len(body) + 8


def save_compression(image_seg_compression: dict, filename) -> int:
    """The driver's last step (encoder/compression/test.py:39-74): final dict -> .rhccq file."""
    h, w = image_seg_compression["shape"]
    idx = np.asarray(image_seg_compression["indices"]).reshape(h, w)
    palette = image_seg_compression["palette"]
    return save_compressed(lossless_compress_optimized(palette, idx, (h, w)), filename)


def save_encoded(palette_u8: np.ndarray, indices: np.ndarray, filename, *, level: int = 9) -> int:
    """Final palette (uint8 [m,3]) and index plane ([h,w]) of the device pipeline -> .rhccq file."""
    h, w = indices.shape
    return save_compressed(lossless_compress_optimized([tuple(int(v) for v in c) for c in palette_u8],
                                                       np.ascontiguousarray(indices), (int(h), int(w)), level=level),
                           filename, level=level)
