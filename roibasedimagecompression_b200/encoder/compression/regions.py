"""Counterpart of /root/reference/encoder/compression/regions.py."""
from __future__ import annotations

from .clustering import compute_clustering_params, cluster_palette_colors_parallel
from .merging import merge_region_components_simple


def region_quantization(regions_components, original_image_height, original_image_width, quality=50,
                        *, as_arrays: bool = False):
    """regions.py:9-70 — flatten, merge onto the full-image canvas, cluster the merged palette."""
    flat = []                                                       # :18-29
    for rc in regions_components:
        if isinstance(rc, dict):
            flat.append(rc)
        elif isinstance(rc, list):
            flat.extend(c for c in rc if isinstance(c, dict))
    merged = merge_region_components_simple(
        flat, (0, 0, original_image_height, original_image_width), as_arrays=True)[0]   # :34-39 (IndexError if empty)
    eps, min_samples, max_cpc = compute_clustering_params(merged["actual_colors"], quality, "lab")   # :52-54
    return [cluster_palette_colors_parallel(quality, merged, eps=eps, min_samples=min_samples,
                                            max_colors_per_cluster=max_cpc, as_arrays=as_arrays)]     # :62-68
