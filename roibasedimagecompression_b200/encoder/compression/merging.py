"""Counterpart of /root/reference/encoder/compression/merging.py (operator level)."""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from . import clustering as _cl


def merge_region_components_simple(region_components, roi_bbox, *, as_arrays: bool = False):
    """merging.py:8-120 — paint component tiles onto one canvas, palette in first-appearance order.

    Components are painted in reversed list order (:52), so the first listed wins overlaps; black
    never paints (:76); a colour takes the next palette slot when the paint sequence first meets it
    (:77-79), even if that pixel is overwritten later.  0 components -> []; 1 component -> a copy
    with ``actual_colors`` ensured (:13-21).
    """
    if not region_components:                                       # :13-14
        return []
    if len(region_components) == 1:                                 # :16-21
        single = region_components[0].copy()
        if "actual_colors" not in single:
            single["actual_colors"] = len(single.get("palette", []))
        return [single]
    be = _cl._be()
    dev = be.device
    minr, minc, maxr, maxc = (int(v) for v in roi_bbox)
    Hc, Wc = maxr - minr, maxc - minc

    def first_pass(components):
        n = len(components)
        comps = np.zeros((n, 8), dtype=np.int32)
        idx_parts, key_parts = [], []
        poff = koff = 0
        for i, seg in enumerate(components):
            sh, sw = (int(v) for v in seg["shape"])
            keys = _cl._rgb_to_keys(seg["palette"])
            idx = np.asarray(seg["indices"], dtype=np.int64).reshape(-1)
            if idx.size != sh * sw:
                raise ValueError(f"component {i}: {idx.size} indices for shape {(sh, sw)}")
            comps[i] = (poff, sh, sw, int(seg["top_left"][0]) - minr, int(seg["top_left"][1]) - minc, koff, len(keys), i)
            idx_parts.append(np.clip(idx, -1, 2 ** 31 - 1).astype(np.int32))
            key_parts.append(keys)
            poff += idx.size
            koff += len(keys)
        d_comps = torch.from_numpy(comps).to(dev)
        d_idx = torch.from_numpy(np.concatenate(idx_parts) if poff else np.zeros(1, np.int32)).to(dev)
        d_keys = torch.from_numpy(np.concatenate(key_parts) if koff else np.zeros(1, np.int32)).to(dev)
        fpos = torch.full((max(koff, 1),), -1, dtype=torch.int32, device=dev)
        be.call("rhccq_comp_pass", be.ptr(d_comps), n, be.ptr(d_idx), Hc, Wc, 0, 0, be.ptr(fpos), 0, 0, be.stream())
        comp_start = torch.from_numpy(np.concatenate([comps[:, 5], [koff]]).astype(np.int32)).to(dev)
        comp_cnt = torch.from_numpy(comps[:, 6].copy()).to(dev)
        grp = torch.tensor([0, n], dtype=torch.int32, device=dev)
        M = ops.merge_level(be, d_keys, fpos, comp_start, comp_cnt, grp, 1, koff + 2, max_entries=max(koff, 1),
                            max_comps=n)
        ops.check_counts("merge_region_components_simple", M["cnt"][:1])
        return n, d_comps, d_idx, M

    n, d_comps, d_idx, M = first_pass(region_components)
    if int(M["present"][0]) < 2:
        # Fewer than two listed components carry palette rows: the entry-level kernel hands a lone component on
        # as it is (what the drivers above it need), but this function still paints a fresh canvas (:27-82) —
        # palette = black + the colours the paint sequence meets, in order.  A last component without pixels
        # and with one colour that appears nowhere makes the kernel take its merging path; it paints nothing
        # and its colour, met by no pixel, takes no palette slot.
        if int(M["present"][0]) == 0:
            pal0 = np.zeros((1, 3), np.uint8)
            idx0 = np.zeros(Hc * Wc, np.int64)
            return [{"top_left": (minr, minc), "shape": (Hc, Wc),
                     "palette": pal0 if as_arrays else [(0, 0, 0)], "indices": idx0 if as_arrays else idx0.tolist(),
                     "indices_dtype": str(np.uint8), "method": "merged", "actual_colors": 1, "encoding": "roi_merged"}]
        ghost = {"top_left": (minr, minc), "shape": (0, 0), "palette": [[1, 1, 1]], "indices": []}
        n, d_comps, d_idx, M = first_pass(list(region_components) + [ghost])
    m = int(M["cnt"][0])
    canvas = torch.zeros((Hc, Wc), dtype=torch.int32, device=dev)
    prio = torch.full((Hc, Wc), 2 ** 31 - 1, dtype=torch.int32, device=dev)
    be.call("rhccq_comp_pass", be.ptr(d_comps), n, be.ptr(d_idx), Hc, Wc, 1, be.ptr(M["map"]), 0, be.ptr(prio), 0,
            be.stream())
    be.call("rhccq_comp_pass", be.ptr(d_comps), n, be.ptr(d_idx), Hc, Wc, 2, be.ptr(M["map"]), 0, be.ptr(prio),
            be.ptr(canvas), be.stream())
    o = int(M["off"][0])
    pal = _cl._keys_to_rgb(M["color"][o:o + m].cpu().numpy())
    indices = canvas.reshape(-1).cpu().numpy().astype(np.int64)
    dt = np.uint8 if m <= 256 else (np.uint16 if m <= 65536 else np.uint32)          # :99-104
    return [{
        "top_left": (minr, minc),
        "shape": (Hc, Wc),
        "palette": pal if as_arrays else [tuple(int(v) for v in c) for c in pal],
        "indices": indices if as_arrays else indices.tolist(),
        "indices_dtype": str(dt),
        "method": "merged",
        "actual_colors": m,
        "encoding": "roi_merged",
    }]
