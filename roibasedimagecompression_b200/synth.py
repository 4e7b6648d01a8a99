"""Deterministic synthetic inputs.

The reference ships no benchmark inputs beyond 25 photographs, and its ROI and
SLIC front ends (encoder/ROI, encoder/subregions) are outside the hot path.
These generators stand in for them: ``synth`` makes an image whose colour
statistics resemble a photograph's (a 64x64 tile holds ~3 000 distinct
colours), ``tile_regions`` makes the region dicts and label maps the stage-1
driver consumes (the shape ``extract_connected_regions_fast`` returns,
/root/reference/encoder/ROI/roi.py:349-358, plus the label map
``enhanced_slic_with_texture`` would return, encoder/subregions/slic.py:99-104).
Host-side numpy only; nothing here is timed.
"""
from __future__ import annotations

import numpy as np


def synth(H: int, W: int, seed: int, sigma: float = 1.0) -> np.ndarray:
    """uint8 [H,W,3]; no pixel is true black (the reference special-cases black)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.empty((H, W, 3), dtype=np.uint8)
    for ch in range(3):
        f = rng.uniform(20.0, 200.0, 4)
        p = rng.uniform(0.0, 2.0 * np.pi, 4)
        v = (128.0
             + 55.0 * np.sin(x / f[0] + p[0]) * np.cos(y / f[1] + p[1])
             + 40.0 * np.sin((x + y) / f[2] + p[2])
             + 20.0 * np.cos((x - y) / f[3] + p[3])
             + rng.normal(0.0, sigma, (H, W)))
        out[..., ch] = np.clip(v, 1, 255).astype(np.uint8)
    return out


def tile_label_map(H: int, W: int, tile: int = 64) -> np.ndarray:
    """int32 [H,W]: id = 1 + ty * ntx + tx for square tiles of ``tile`` pixels."""
    ntx = -(-W // tile)
    ty = np.arange(H) // tile
    tx = np.arange(W) // tile
    return (1 + ty[:, None] * ntx + tx[None, :]).astype(np.int32)


def tile_is_roi(H: int, W: int, tile: int = 64) -> np.ndarray:
    """bool [H,W]: the checkerboard "ROI" — tiles with (tx + ty) even."""
    ty = np.arange(H) // tile
    tx = np.arange(W) // tile
    return ((ty[:, None] + tx[None, :]) % 2) == 0


def _region_from_mask(mask: np.ndarray, labels: np.ndarray) -> dict:
    rows = np.flatnonzero(mask.any(axis=1))
    cols = np.flatnonzero(mask.any(axis=0))
    minr, maxr = int(rows[0]), int(rows[-1]) + 1
    minc, maxc = int(cols[0]), int(cols[-1]) + 1
    bm = mask[minr:maxr, minc:maxc]
    seg = np.where(bm, labels[minr:maxr, minc:maxc], 0).astype(np.int32)
    return {"bbox": (minr, minc, maxr, maxc), "bbox_mask": bm, "segments": seg,
            "area": int(mask.sum())}


def tile_regions(H: int, W: int, tile: int = 64):
    """(roi_regions, nonroi_regions): one region per class, tiles as segments."""
    labels = tile_label_map(H, W, tile)
    roi = tile_is_roi(H, W, tile)
    out = []
    for m in (roi, ~roi):
        out.append([_region_from_mask(m, labels)] if m.any() else [])
    return out[0], out[1]


def pixel_features(image_rgb: np.ndarray, w_xy: float = 1.0) -> np.ndarray:
    """float32 [H*W,5] points (w_xy*x, w_xy*y, R, G, B) in raster order."""
    H, W, _ = image_rgb.shape
    y, x = np.mgrid[0:H, 0:W]
    return np.concatenate([(w_xy * x)[..., None], (w_xy * y)[..., None], image_rgb],
                          axis=2).reshape(-1, 5).astype(np.float32)
