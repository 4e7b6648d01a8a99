"""Container writer / reader and the decoder gather + metrics kernels (N4).

The writer must produce the reference's bytes for the same input (golden written by the reference's own
compression.py, tests/golden/make_golden.py container); the reader must read a file the reference shipped."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import rhccq_oracle as O
from roibasedimagecompression_b200.encoder import compression as C
from roibasedimagecompression_b200.decoder.uncompression import uncompression as U


def test_writer_is_byte_identical_to_the_reference(tmp_path):
    g = golden("pipeline_small.npz")
    pal, idx, shape = g["palette"], g["indices"], tuple(int(v) for v in g["shape"])
    fn = tmp_path / "mine.rhccq"
    size = C.save_encoded(pal, idx.reshape(shape), fn)
    want = open(os.path.join(GOLDEN, "container_small.rhccq"), "rb").read()
    got = open(fn, "rb").read()
    assert len(got) == len(want) and size == len(got) - 1     # the reference reports body + 8 for a 9-byte header (compression.py:142)
    assert got == want
    # the driver-level entry (encoder/compression/test.py:39-74) on the reference's dict form
    fn2 = tmp_path / "mine2.rhccq"
    C.save_compression({"shape": shape, "palette": pal.tolist(), "indices": idx.astype(int).tolist()}, fn2)
    assert open(fn2, "rb").read() == want


def test_writer_at_another_zlib_level_is_read_back(tmp_path):
    """``level`` trades size for host time; the reader (and the reference's: it only calls zlib.decompress) gets the
    same palette and indices from any level, only level 9 gives the reference's bytes."""
    g = golden("pipeline_small.npz")
    pal, idx, shape = g["palette"], g["indices"], tuple(int(v) for v in g["shape"])
    want = open(os.path.join(GOLDEN, "container_small.rhccq"), "rb").read()
    for level in (1, 6):
        fn = tmp_path / f"l{level}.rhccq"
        C.save_encoded(pal, idx.reshape(shape), fn, level=level)
        assert open(fn, "rb").read() != want
        p2, i2, s2 = U.lossless_decompress(U.load_compressed(fn))
        assert np.array_equal(p2, pal) and np.array_equal(np.asarray(i2).reshape(shape), idx.reshape(shape)) and tuple(s2) == shape


def test_reader_reads_a_file_shipped_by_the_reference():
    d = U.load_compressed(os.path.join(GOLDEN, "reference_Lenna_compressed.rhccq"))
    assert set(d) == {"s", "l", "p", "i", "d"} and tuple(d["s"]) == (512, 512) and d["l"] == 147 and d["d"] == "uint8"
    pal, idx, shape = U.lossless_decompress(d)
    assert pal.shape == (147, 3) and idx.dtype == np.uint8 and idx.size == 512 * 512
    assert pal[0].tolist() == [0, 0, 0] and not (idx == 0).any()          # black row 0 is never used


def test_reader_refuses_foreign_pickles(tmp_path):
    import pickle, struct, zlib
    body = zlib.compress(pickle.dumps({"s": os.getcwd}, protocol=5), 9)   # a global that is not numpy's
    fn = tmp_path / "evil.rhccq"
    fn.write_bytes(b"RHCCQ" + struct.pack("<I", len(body)) + body)
    with pytest.raises(Exception):
        U.load_compressed(fn)
    fn.write_bytes(b"NOPE!" + struct.pack("<I", 0))
    with pytest.raises(ValueError):
        U.load_compressed(fn)


def test_decode_and_metrics_kernels(backend, tmp_path):
    U._BACKEND = backend if backend.device.type == "cpu" else None
    try:
        g = golden("pipeline_small.npz")
        img, pal, idx, shape = g["image"], g["palette"], g["indices"], tuple(int(v) for v in g["shape"])
        fn = tmp_path / "rt.rhccq"
        C.save_encoded(pal, idx.reshape(shape), fn)
        p2, i2, s2 = U.lossless_decompress(U.load_compressed(fn))
        rec = U.decompress_color_quantization(p2, i2, s2)
        assert np.array_equal(rec, pal[idx.reshape(shape)])
        m = U.quality_metrics(rec, img)
        assert abs(m["psnr"] - float(g["psnr"])) < 1e-9 and abs(m["psnr"] - O.psnr(rec, img)) < 1e-9
        assert abs(m["mae"] - np.abs(rec.astype(int) - img.astype(int)).mean()) < 1e-12
        for dt in (np.uint8, np.uint16, np.uint32):
            r = U.decompress_color_quantization(pal, idx.astype(dt), shape)
            assert np.array_equal(r, rec)
        with pytest.raises(Exception):
            U.decompress_color_quantization(pal[:3], idx, shape)          # index outside the palette
        assert U.quality_metrics(img, img)["psnr"] == float("inf")
    finally:
        U._BACKEND = None
