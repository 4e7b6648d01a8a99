"""Counterpart of /root/reference/encoder/compression/image.py (quantize_image) and of
optimize_compressed_dtype (/root/reference/encoder/compression/compression.py:326-413)."""
from __future__ import annotations

import numpy as np

from .clustering import compute_clustering_params, cluster_palette_colors_parallel
from .merging import merge_region_components_simple


def optimize_compressed_dtype(comp: dict) -> dict:
    """compression.py:326-413: smallest unsigned dtype that holds max(indices); values unchanged."""
    if "indices" not in comp:
        return comp
    idx = np.asarray(comp["indices"])
    mx = int(idx.max()) if idx.size else 0
    name = "uint8" if mx < 256 else ("uint16" if mx < 65536 else "uint32")
    out = comp.copy()
    out["indices_dtype"] = name
    out["indices_optimized"] = True
    out["actual_colors"] = len(comp["palette"])
    return out


def quantize_image(image_components, original_image_height, original_image_width, quality=100,
                   *, as_arrays: bool = False):
    """image.py:243-289 — merge ROI + non-ROI, cluster, pick the index dtype."""
    merged = merge_region_components_simple(
        list(image_components), (0, 0, original_image_height, original_image_width), as_arrays=True)[0]
    eps, min_samples, max_cpc = compute_clustering_params(merged["actual_colors"], quality, "lab")
    comp = cluster_palette_colors_parallel(quality, merged, eps=eps, min_samples=min_samples,
                                           max_colors_per_cluster=max_cpc, as_arrays=as_arrays)
    return optimize_compressed_dtype(comp)
