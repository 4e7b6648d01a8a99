"""Segment label maps from the pixel-feature DBSCAN, in the slot the reference fills with SLIC.

The reference calls ``enhanced_slic_with_texture(bbox_region, bbox_mask, n_segments)``
(/root/reference/encoder/subregions/slic.py:41-104, call site
/root/reference/encoder/compression/subregions.py:160) and gets ``(segments int32 [h, w], texture_map)``:
scikit-image SLIC on a down-scaled copy, 0 outside the mask, and a texture map that is all zeros
(slic.py:49, the Gabor block is commented out).  scikit-image is not part of this path; what the north star
puts in this slot is the DBSCAN of the (x, y, R, G, B) pixel features.  This is a DIFFERENT segmentation
algorithm, so there is no bit-exact oracle for the label map itself; what is pinned is
everything downstream of it: the quantiser run on these label maps is bit-exact against the oracle run on the
same label maps (tests/test_segmenter.py), and the decoded PSNR is gated against the tile segmentation.

DBSCAN leaves noise pixels and, on natural images, thousands of tiny clusters next to a few large ones; a
segmenter has to give every masked pixel a segment and should stay near the requested ``n_segments``.  The
policy (the reference defines none):

* clusters of at least ``min_size`` masked pixels become segments of their own, numbered by first pixel;
* every other masked pixel (noise, small clusters) goes to the residual segment of its square tile of side
  ``ceil(sqrt(masked area / n_segments))`` — so the residual contributes about ``n_segments`` segments.

Labels are 1-based and dense; 0 = outside the mask, as the reference's SLIC call returns them.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ... import dbscan as D
from ..compression import clustering as _cl


class DbscanSegmenter:
    """``segmenter(bbox_image, bbox_mask) -> int32 [h, w]`` for ``subregion_quantization(..., segmenter=...)``."""

    def __init__(self, be=None, eps: float = 3.0, min_pts: int = 8, n_segments: int = 100, min_size: int | None = None):
        self.be = be if be is not None else _cl._be()               # the CUDA library (fails loudly without it)
        self.eps, self.min_pts, self.n_segments, self.min_size = float(eps), int(min_pts), int(n_segments), min_size

    def __call__(self, image, mask, n_segments: int | None = None):
        be = self.be
        img = torch.as_tensor(np.ascontiguousarray(image, dtype=np.uint8)).to(be.device)
        h, w, _ = img.shape
        m = torch.as_tensor(np.ascontiguousarray(mask, dtype=bool)).to(be.device).view(-1)
        area = int(m.sum().item())
        out = torch.zeros(h * w, dtype=torch.int32, device=be.device)
        if area == 0:
            return out.view(h, w).cpu().numpy()
        k = max(int(n_segments if n_segments is not None else self.n_segments), 1)
        lab, _core = D.dbscan_image(be, img, self.eps, self.min_pts)             # int32 [h, w], -1 = noise
        lab = lab.view(-1).to(torch.int64)
        n_cl = int(lab.max().item()) + 1
        min_size = self.min_size if self.min_size is not None else max(16, area // (4 * k))
        side = max(int(math.ceil(math.sqrt(area / k))), 1)
        ys = torch.arange(h, device=be.device).view(h, 1).expand(h, w).reshape(-1)
        xs = torch.arange(w, device=be.device).view(1, w).expand(h, w).reshape(-1)
        tiles_x = (w + side - 1) // side
        tile = (ys // side) * tiles_x + xs // side                                # residual segment key
        keep = torch.zeros_like(m)
        if n_cl > 0:
            inside = m & (lab >= 0)
            size = torch.bincount(lab[inside], minlength=n_cl)                    # masked pixels per cluster
            keep = inside & (size[lab.clamp(min=0)] >= min_size)
        # one key space: kept clusters first (they are numbered by lowest pixel index already), then tiles
        key = torch.where(keep, lab, n_cl + tile)
        key = torch.where(m, key, torch.full_like(key, -1))
        uniq, inv = torch.unique(key[m], return_inverse=True)                     # sorted: dense 1-based ids
        out[m] = (inv + 1).to(torch.int32)
        return out.view(h, w).cpu().numpy()


def enhanced_slic_with_texture(image, mask, n_segments=100, compactness=10, *, eps: float = 3.0, min_pts: int = 8, be=None):
    """Same call and return contract as the reference's function of this name (slic.py:41): ``(segments,
    texture_map)`` with ``segments`` int32 [h, w], 0 outside ``mask``.  ``compactness`` is accepted and unused
    (it is a SLIC parameter); the texture map is all zeros, as in the reference (slic.py:49)."""
    seg = DbscanSegmenter(be, eps=eps, min_pts=min_pts, n_segments=n_segments)(image, mask)
    return seg, np.zeros(seg.shape, dtype=np.float64)
