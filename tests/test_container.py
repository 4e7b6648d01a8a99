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


def test_device_zlib_streams_decompress_to_the_index_bytes(backend):
    """csrc/rhccq_deflate.cu: whatever the plane holds — runs, rows repeating the row above, noise with bytes on both
    sides of the 8 / 9-bit literal codes, one or two bytes per index, sizes that are not multiples of the 4 KiB
    chunk, planes shorter than a match — zlib.decompress returns the bytes (and checks the Adler-32)."""
    import zlib
    from roibasedimagecompression_b200.encoder.compression import device_deflate as DD
    rng = np.random.default_rng(8)
    cases = []
    g = golden("pipeline_small.npz")
    shape = tuple(int(v) for v in g["shape"])
    cases.append((g["indices"].astype(np.uint8).reshape(1, -1), 1, shape[1]))
    cases.append((np.ascontiguousarray(g["indices"].astype(np.uint16).reshape(1, -1)).view(np.uint8), 2, 2 * shape[1]))
    cases.append((rng.integers(0, 256, (3, 9000), dtype=np.uint8), 1, 300))                    # incompressible
    smooth = np.repeat(np.repeat(rng.integers(0, 200, (2, 20, 30), dtype=np.uint8), 7, axis=1), 11, axis=2)
    cases.append((smooth.reshape(2, -1), 1, smooth.shape[2]))                                  # runs and repeated rows
    cases.append((np.zeros((1, 70000), np.uint8), 1, 40000))                                   # row longer than the window
    cases.append((np.full((2, 2), 255, np.uint8), 1, 2))
    cases.append((np.arange(1, dtype=np.uint8).reshape(1, 1), 1, 1))
    wide = (np.arange(5000) // 37 % 700).astype(np.uint16)
    cases.append((np.ascontiguousarray(wide.reshape(1, -1)).view(np.uint8), 2, 200))
    for planes, elem, row in cases:
        planes = np.ascontiguousarray(planes)
        got = DD.zlib_streams(backend, torch.from_numpy(planes).to(backend.device), elem, row)
        assert len(got) == planes.shape[0]
        for f, st in enumerate(got):
            assert zlib.decompress(st) == planes[f].tobytes(), (planes.shape, elem, row)
    empty = DD.zlib_streams(backend, torch.zeros((2, 0), dtype=torch.uint8, device=backend.device), 1, 1)
    assert empty == [b"\x78\x01\x03\x00\x00\x00\x00\x01"] * 2 and zlib.decompress(empty[0]) == b""


def test_device_container_is_read_by_the_reader(backend, tmp_path):
    """save_batch: index streams from the device inside the reference's package; the reader returns the pipeline's
    palette and indices."""
    from roibasedimagecompression_b200 import pipeline
    from roibasedimagecompression_b200.encoder.compression import device_deflate as DD
    from roibasedimagecompression_b200.synth import synth
    H, W = 64, 96
    imgs = np.stack([synth(H, W, 5 + i) for i in range(2)])
    tab, lab = pipeline.table_from_tiles(2, H, W, 32)
    labs = np.ascontiguousarray(np.broadcast_to(lab, (2, 2, H, W)))
    res = pipeline.encode_batch(backend, torch.from_numpy(imgs).to(backend.device), torch.from_numpy(labs).to(backend.device), tab)
    pipeline.finish_checks(res)
    names = [tmp_path / f"f{b}.rhccq" for b in range(2)]
    sizes = DD.save_batch(backend, res, names)
    for b in range(2):
        assert sizes[b] == os.path.getsize(names[b])
        p2, i2, s2 = U.lossless_decompress(U.load_compressed(names[b]))
        assert tuple(s2) == (H, W)
        assert np.array_equal(p2, res.palette(b))
        assert np.array_equal(np.asarray(i2).reshape(H, W), res.index_image(b))


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
