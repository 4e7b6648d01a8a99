"""Counterpart of /root/reference/encoder/compression/clustering.py (operator level).

``get_all_unique_colors``, ``compute_clustering_params`` and
``cluster_palette_colors_parallel`` keep the reference's signatures and return
dicts with the reference's keys; ``palette`` / ``indices`` are Python lists as
in the reference unless ``as_arrays=True`` (numpy arrays, for callers that
would turn them straight back into arrays, e.g. encoder/compression/test.py:48-54).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ... import ops
from ..._lib import lib, RhccqError

_BACKEND = None          # tests bind the host-emulation build here; the product uses lib()


def _be():
    return _BACKEND if _BACKEND is not None else lib()


def _keys_to_rgb(keys: np.ndarray) -> np.ndarray:
    k = keys.astype(np.uint32)
    return np.stack([(k >> 16) & 255, (k >> 8) & 255, k & 255], axis=-1).astype(np.uint8)


def _rgb_to_keys(pal) -> np.ndarray:
    a = np.asarray(pal, dtype=np.int64).reshape(-1, 3)
    if a.size and (a.min() < 0 or a.max() > 255):
        raise ValueError("palette entries must be 0..255")
    return ((a[:, 0] << 16) | (a[:, 1] << 8) | a[:, 2]).astype(np.int32)


def get_all_unique_colors(region_image, top_left_coords, *, as_arrays: bool = False):
    """clustering.py:4-103 — every distinct colour of the crop, lexicographic, and the row of each pixel."""
    if region_image is None or getattr(region_image, "size", 0) == 0:        # :9-10
        return None
    be = _be()
    img = np.ascontiguousarray(region_image, dtype=np.uint8)
    h, w, _ = img.shape
    total = h * w
    d_img = torch.from_numpy(img).to(be.device).reshape(1, h, w, 3)
    crops = torch.tensor([[0, 0, 0, h, w, 0, 0, 0]], dtype=torch.int32, device=be.device)
    pal_off = torch.zeros(1, dtype=torch.int32, device=be.device)
    keys, cnt, plane = ops.unique_index(be, d_img, None, crops, pal_off, total, idx_bytes=4, max_valid=total)
    ops.check_counts("get_all_unique_colors", cnt[:1])
    n = int(cnt[0])
    palette = _keys_to_rgb(keys[:n].cpu().numpy())
    indices = plane.reshape(-1).cpu().numpy().astype(np.int64)
    bytes_per_index = 1 if n <= 256 else 2                                    # :57-62
    compressed = n * 3 + total * bytes_per_index + 50                         # :65-69
    return {
        "method": "exact_colors",
        "top_left": top_left_coords,
        "shape": (h, w),
        "palette": palette if as_arrays else palette.tolist(),
        "indices": indices if as_arrays else indices.tolist(),
        "max_colors": n,
        "actual_colors": n,
        "index_dtype": str(np.uint8 if n <= 256 else np.uint16),
        "original_size": total * 3,
        "compressed_size": compressed,
        "compression_ratio": (total * 3) / compressed if compressed > 0 else 0,
        "mse": 0.0,
        "psnr": float("inf"),
        "encoding": "exact",
    }


def compute_clustering_params(n_colors, quality, color_space="rgb"):
    """clustering.py:108-135 (``color_space`` is ignored there as well)."""
    eps = 128 - 1.28 * quality
    max_colors_per_cluster = math.ceil((-(quality / 100) * n_colors + n_colors) / quality)
    if eps == 0:
        eps = 1
    if max_colors_per_cluster == 0:
        max_colors_per_cluster = 1
    min_samples = 1
    return eps, min_samples, max_colors_per_cluster


def _cluster_device(be, keys_np: np.ndarray, eps: float, max_cpc: int, leaf_override=None, quality: float = 0.0):
    """Run a3' + a3/a4 + means for one palette; returns (labels, leaf, n_leaves, new_keys) as numpy."""
    n = len(keys_np)
    dev = be.device
    keys = torch.from_numpy(np.ascontiguousarray(keys_np, dtype=np.int32)).to(dev)
    off = torch.zeros(1, dtype=torch.int32, device=dev)
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    thr, tie = ops.eps_threshold(eps)
    slots = be.cdll.rhccq_palette_dbscan_slots(thr, n)
    labels, ncl = ops.palette_dbscan(
        be, keys, off, cnt, torch.tensor([thr], dtype=torch.int32, device=dev),
        torch.tensor([tie], dtype=torch.int32, device=dev), torch.tensor([float(eps)], dtype=torch.float64, device=dev),
        max_rows=n, max_slots=slots)
    if n >= 10000:                                                   # clustering.py:207-218 may apply
        ops.palette_minibatch(be, keys, off, cnt, torch.tensor([float(quality)], dtype=torch.float64, device=dev),
                              labels, ncl, max_rows=n)
    ops.check_counts("cluster_palette_colors_parallel (DBSCAN)", ncl[:1])
    if leaf_override is None:
        mc = torch.tensor([int(max_cpc)], dtype=torch.int32, device=dev)
        leaf, nl = ops.palette_split(be, keys, off, cnt, labels, ncl, mc, max_rows=n)
        ops.check_counts("cluster_palette_colors_parallel (split)", nl[:1])
    else:
        lo = np.asarray(leaf_override, dtype=np.int32)
        leaf = torch.from_numpy(lo).to(dev)
        nl = torch.tensor([int(lo.max()) + 1 if lo.size else 0], dtype=torch.int32, device=dev)
    new_keys = ops.palette_finish(be, keys, off, cnt, leaf, nl, max_rows=n)
    m = int(nl[0])
    return labels.cpu().numpy(), leaf.cpu().numpy(), m, new_keys[:m].cpu().numpy()


def cluster_palette_colors_parallel(quality, compressed_data, eps=10.0, min_samples=2,
                                    max_colors_per_cluster=5, num_workers=None, *,
                                    as_arrays: bool = False, leaf_override=None):
    """clustering.py:160-437.

    ``min_samples`` must be 1 — every call site of the reference passes 1
    (subregions.py:447, regions.py:66, image.py:278); general DBSCAN lives in
    the point-cloud path.  Large clusters are consumed in submission order; the
    reference consumes them in thread-completion order (:458), its one
    nondeterministic step.  ``leaf_override`` (tests only) supplies the new
    palette row of every old row instead of the K-Means split, so that
    everything downstream of K-Means can be compared with the reference bit
    for bit when scikit-learn's own assignment is injected.
    """
    if min_samples != 1:
        raise NotImplementedError("palette DBSCAN is built for min_samples=1, the value every reference "
                                  "call site passes")
    be = _be()
    palette = np.asarray(compressed_data["palette"], dtype=np.uint8).reshape(-1, 3)   # :171
    indices = np.asarray(compressed_data["indices"])                                   # :172
    h, w = compressed_data["shape"]
    n_orig = len(palette)
    keys = _rgb_to_keys(palette)
    if not np.any(keys != 0):                                                          # :197-199
        return compressed_data
    labels, leaf, m, new_keys = _cluster_device(be, keys, float(eps), int(max_colors_per_cluster), leaf_override,
                                                quality=float(quality))
    new_palette = _keys_to_rgb(new_keys)
    if len(np.unique(keys)) != n_orig and leaf_override is None:
        # A palette handed on unmerged (merging.py:16-21) can hold a colour twice.  Rows of clusters that were
        # split go back to palette rows through find_color_index (:803-808), which finds the FIRST row of a
        # colour: that row takes the entry of the last split holding the colour (equal colours always share a
        # split), every later duplicate row keeps the table's initial 0 (:373).
        sizes = np.bincount(labels[labels >= 0], minlength=max(int(labels.max()) + 1, 1))
        in_split = (labels >= 0) & (sizes[np.maximum(labels, 0)] > int(max_colors_per_cluster)) & (keys != 0)
        _, first = np.unique(keys, return_index=True)
        dup = np.ones(n_orig, dtype=bool)
        dup[first] = False
        leaf = np.where(in_split & dup, 0, leaf)
    lut = leaf.astype(np.uint16)                                                       # :373 (uint16 table)
    new_indices = lut[indices.astype(np.int64).ravel()].astype(np.int64)              # :377
    total = h * w
    original_size = compressed_data.get("original_size", total * 3)
    bpi = 1 if m <= 256 else 2
    new_size = m * 3 + total * bpi + 100
    return {
        "method": "clustered_colors",
        "top_left": compressed_data["top_left"],
        "shape": (h, w),
        "palette": new_palette if as_arrays else new_palette.tolist(),
        "indices": new_indices if as_arrays else new_indices.tolist(),
        "original_unique_colors": n_orig,
        "compressed_colors": m,
        "index_dtype": "uint8" if m <= 256 else "uint16",
        "original_size": original_size,
        "compressed_size": new_size,
        "compression_ratio": original_size / new_size if new_size > 0 else 0,
        "mse": 0.0,
        "psnr": float("inf"),
        "clustering_params": {"eps": eps, "min_samples": min_samples,
                              "max_colors_per_cluster": max_colors_per_cluster},
        "encoding": "dbscan_clustered",
        "black_preserved": True,
        "parallel_processed": True,
    }


def dbscan_palette_labels(palette_rgb, eps: float) -> np.ndarray:
    """Labels of ``DBSCAN(eps/255, min_samples=1).fit_predict(palette/255.0)`` (clustering.py:204-235).

    Not a reference function: the third-party operator one level below it, exposed
    for parity tests and for callers that only need the components.  Black rows
    are clustered like any other colour here (the caller removes them first, as
    clustering.py:185-192 does).
    """
    be = _be()
    keys = _rgb_to_keys(palette_rgb)
    if np.any(keys == 0):
        raise ValueError("remove black rows first (clustering.py:185-192)")
    labels, _, _, _ = _cluster_device(be, keys, float(eps), len(keys) + 1)
    return labels.astype(np.int64)
