"""Counterpart of /root/reference/encoder/compression/subregions.py (stage-1 driver)."""
from __future__ import annotations

import numpy as np
import torch

from ... import ops, pipeline
from . import clustering as _cl


def subregion_quantization(image_rgb, subregions, quality=10, subregion_type=None, debug=False,
                           *, segmenter=None, as_arrays: bool = False):
    """subregions.py:90-683 — per region: segments -> unique colours -> cluster -> merge.

    The reference obtains the segment label map of every region from scikit-image SLIC
    (``enhanced_slic_with_texture``, subregions.py:160).  That front end is outside this
    path: each region dict must carry the label map as
    ``region['segments']`` (int32, bbox-sized, 0 = outside), or ``segmenter(bbox_image,
    bbox_mask) -> labels`` must be given.  Returns the reference's list (one entry per region)
    of lists of component dicts: ``[merged]`` for a region of several segments, the single
    clustered component otherwise, ``[]`` for a region without segments.
    """
    be = _cl._be()
    img = np.ascontiguousarray(image_rgb, dtype=np.uint8)
    H, W, _ = img.shape
    regions = []
    for region in subregions:
        if "segments" not in region:
            if segmenter is None:
                raise ValueError("region['segments'] is missing and no segmenter was given: the SLIC front "
                                 "end (encoder/subregions/slic.py) is not part of the hot path")
            minr, minc, maxr, maxc = region["bbox"]
            region = dict(region)
            region["segments"] = np.asarray(segmenter(img[minr:maxr, minc:maxc], region["bbox_mask"]), np.int32)
        regions.append(region)
    tab, lab = pipeline.table_from_regions((H, W), [regions], (quality,))
    out = [[] for _ in regions]
    if tab.P == 0:
        return out
    dev = be.device
    d_img = torch.from_numpy(img[None]).to(dev)
    d_lab = torch.from_numpy(lab).to(dev)
    st = pipeline.stage1(be, d_img, d_lab, tab)
    A = st["A"]
    a_off, a_cnt, a_present = (A[k].cpu().numpy() for k in ("off", "cnt", "present"))
    a_color = A["color"].cpu().numpy()
    a_map = A["map"].cpu().numpy()
    plane = st["plane"][0, 0].cpu().numpy()
    plane = (plane.view(np.uint16) if st["idx_bytes"] == 2 else plane.view(np.uint32)).astype(np.int64)
    label = lab[0, 0]
    nl1, ent_off = st["nl1"], st["ent_off_h"]
    pal_cnt_h = st["pal_cnt"].cpu().numpy()
    crops = tab.crops
    for r, region in enumerate(regions):
        segs = np.flatnonzero(tab.seg_region == r)
        if segs.size == 0:
            continue
        minr, minc, maxr, maxc = (int(v) for v in region["bbox"])
        pal = _cl._keys_to_rgb(a_color[a_off[r]:a_off[r] + a_cnt[r]])
        if segs.size == 1:                                          # subregions.py:679: the clustered segment itself
            p = int(segs[0])
            _, r0, c0, h, w, sid, _, _ = (int(v) for v in crops[p])
            inside = label[r0:r0 + h, c0:c0 + w] == sid
            idx = np.where(inside, plane[r0:r0 + h, c0:c0 + w], 0).reshape(-1)
            m = len(pal)
            total = h * w
            n_orig = int(pal_cnt_h[p])
            eps, ms, mcpc = _cl.compute_clustering_params(n_orig, quality, "lab")
            new_size = m * 3 + total * (1 if m <= 256 else 2) + 100
            out[r] = [{
                "method": "clustered_colors", "top_left": (r0, c0), "shape": (h, w),
                "palette": pal if as_arrays else pal.tolist(),
                "indices": idx if as_arrays else idx.tolist(),
                "original_unique_colors": n_orig, "compressed_colors": m,
                "index_dtype": "uint8" if m <= 256 else "uint16",
                "original_size": total * 3, "compressed_size": new_size,
                "compression_ratio": total * 3 / new_size, "mse": 0.0, "psnr": float("inf"),
                "clustering_params": {"eps": eps, "min_samples": ms, "max_colors_per_cluster": mcpc},
                "encoding": "dbscan_clustered", "black_preserved": True, "parallel_processed": True,
            }]
            continue
        canvas = np.zeros((maxr - minr, maxc - minc), dtype=np.int64)
        for p in segs:                                              # masks are disjoint: order is immaterial
            _, r0, c0, h, w, sid, _, _ = (int(v) for v in crops[p])
            inside = label[r0:r0 + h, c0:c0 + w] == sid
            sub = canvas[r0 - minr:r0 - minr + h, c0 - minc:c0 - minc + w]
            sub[inside] = a_map[ent_off[p] + plane[r0:r0 + h, c0:c0 + w][inside]]
        m = len(pal)
        dt = np.uint8 if m <= 256 else (np.uint16 if m <= 65536 else np.uint32)
        out[r] = [{
            "top_left": (minr, minc), "shape": (maxr - minr, maxc - minc),
            "palette": pal if as_arrays else [tuple(int(v) for v in c) for c in pal],
            "indices": canvas.reshape(-1) if as_arrays else canvas.reshape(-1).tolist(),
            "indices_dtype": str(dt), "method": "merged", "actual_colors": m, "encoding": "roi_merged",
        }]
    return out
