"""Tensor-level wrappers of the C-ABI ops (include/rhccq.h).

Every function takes the backend first, then tensors that live on the
backend's device, and returns freshly allocated tensors.  No arithmetic of the
hot path happens here: torch provides memory and streams only.
"""
from __future__ import annotations

import math

import torch

from ._lib import Backend, RhccqError

I32 = torch.int32


def eps_threshold(eps: float):
    """Integer form of scikit-learn's radius predicate on 0..255 colours.

    The reference runs ``DBSCAN(eps/255).fit_predict(palette/255.0)``
    (encoder/compression/clustering.py:205,233-235); for integer colours the
    KD-tree's ``sum(((a-b)/255)^2) <= (eps/255)^2`` is ``d2 <= floor(eps^2)``
    unless ``eps^2`` is an integer, where ``d2 == eps^2`` is a float64 tie the
    kernel evaluates the way the KD-tree does.  Returns (thr, tie).
    """
    e2 = float(eps) * float(eps)
    tie = abs(e2 - round(e2)) < 1e-9
    thr = int(round(e2)) if tie else int(math.floor(e2 + 1e-9))
    return thr, int(tie)


def _as_dev(be: Backend, x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(device=be.device, dtype=dtype).contiguous()
    import numpy as np
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dtype).to(be.device).contiguous()
    return torch.as_tensor(x, dtype=dtype).to(be.device).contiguous()


def unique_index(be: Backend, img, seg, crops, pal_off, pal_capacity: int, *, idx_bytes: int = 2,
                 repaint_black: bool = False, max_valid: int, n_classes: int = 1):
    """a1 — rhccq_unique_index.  Returns (pal_keys, pal_cnt, index_plane [K,B,H,W])."""
    B, H, W, _ = img.shape
    P = crops.shape[0]
    pal_keys = be.empty((max(pal_capacity, 1),), torch.int32)      # uint32 keys, viewed as int32 storage
    pal_cnt = be.empty((max(P, 1),), I32)
    plane = be.zeros((n_classes, B, H, W), torch.int16 if idx_bytes == 2 else torch.int32)
    need = be.cdll.rhccq_unique_index_workspace_bytes(max_valid)
    ws, ws_bytes = be.workspace(need, P)
    be.call("rhccq_unique_index", be.ptr(img), be.ptr(seg), B, H, W, be.ptr(crops), P, be.ptr(pal_off),
            be.ptr(pal_keys), be.ptr(pal_cnt), be.ptr(plane), idx_bytes, int(repaint_black), int(max_valid),
            be.ptr(ws), ws_bytes, be.stream())
    return pal_keys, pal_cnt, plane


def cluster_params(be: Backend, n_colors, quality):
    """a2 on device-resident counts — rhccq_cluster_params.  Returns max_cpc (int32)."""
    n = n_colors.numel()
    out = be.empty((max(n, 1),), I32)
    be.call("rhccq_cluster_params", be.ptr(n_colors), be.ptr(quality), n, be.ptr(out), be.stream())
    return out


def palette_dbscan(be: Backend, pal_keys, pal_off, pal_cnt, thr, tie, eps, *, max_rows: int, max_slots: int):
    """a3' — rhccq_palette_dbscan.  Returns (labels, n_clusters)."""
    P = pal_cnt.numel()
    labels = be.empty((pal_keys.numel(),), I32)
    ncl = be.empty((max(P, 1),), I32)
    need = be.cdll.rhccq_palette_dbscan_workspace_bytes(max_rows, max_slots)
    ws, ws_bytes = be.workspace(need, P)
    be.call("rhccq_palette_dbscan", be.ptr(pal_keys), be.ptr(pal_off), be.ptr(pal_cnt), be.ptr(thr), be.ptr(tie),
            be.ptr(eps), P, be.ptr(labels), be.ptr(ncl), int(max_rows), int(max_slots), be.ptr(ws), ws_bytes,
            be.stream())
    return labels, ncl


def palette_minibatch(be: Backend, pal_keys, pal_off, pal_cnt, quality, labels, n_clusters, *, max_rows: int):
    """a3, >= 10 000 colours — rhccq_palette_minibatch; labels / n_clusters of the palettes that
    rhccq_palette_dbscan marked with status -4 are filled in place."""
    P = pal_cnt.numel()
    ws_bytes = int(be.cdll.rhccq_palette_minibatch_workspace_bytes(int(max_rows), int(P)))
    ws = be.empty((max(ws_bytes, 1),), torch.uint8)
    be.call("rhccq_palette_minibatch", be.ptr(pal_keys), be.ptr(pal_off), be.ptr(pal_cnt), be.ptr(quality), P,
            be.ptr(labels), be.ptr(n_clusters), int(max_rows), be.ptr(ws), ws_bytes, be.stream(), launches=2)


def palette_split(be: Backend, pal_keys, pal_off, pal_cnt, labels, n_clusters, max_cpc, *, max_rows: int):
    """a3/a4 — rhccq_palette_split.  Returns (leaf, n_leaves)."""
    P = pal_cnt.numel()
    leaf = be.empty((pal_keys.numel(),), I32)
    nl = be.empty((max(P, 1),), I32)
    rng = be.rng_table(be.cdll.rhccq_kmeans_rng_need(int(max_rows)))
    need = be.cdll.rhccq_palette_split_workspace_bytes(max_rows)
    ws, ws_bytes = be.workspace(need, P)
    be.call("rhccq_palette_split", be.ptr(pal_keys), be.ptr(pal_off), be.ptr(pal_cnt), P, be.ptr(labels),
            be.ptr(n_clusters), be.ptr(max_cpc), be.ptr(rng), rng.numel(), be.ptr(leaf), be.ptr(nl), int(max_rows),
            be.ptr(ws), ws_bytes, be.stream())
    return leaf, nl


def kmeans_labels(be: Backend, colors, k: int):
    """``KMeans(n_clusters=k, random_state=42, n_init='auto').fit_predict(colors.astype(float))`` — the
    third-party operator the reference calls at clustering.py:751-752 — through rhccq_palette_split with
    max_cpc = -k.  Returns int64 labels; labels are ranks among the non-empty clusters (identical to
    scikit-learn's unless a cluster ends empty, which scikit-learn reports with a ConvergenceWarning)."""
    import numpy as np
    c = np.ascontiguousarray(np.asarray(colors).astype(np.uint8).reshape(-1, 3)).astype(np.int64)
    n = int(c.shape[0])
    if not 1 <= int(k) <= n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    if (c == 0).all(axis=1).any():
        raise NotImplementedError("black rows never reach K-Means in the reference (clustering.py:185-192)")
    keys = _as_dev(be, ((c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]).astype(np.int32), I32)
    off = _as_dev(be, np.zeros(1, np.int32), I32)
    cnt = _as_dev(be, np.full(1, n, np.int32), I32)
    lab = be.zeros((n,), I32)
    ncl = _as_dev(be, np.ones(1, np.int32), I32)
    mc = _as_dev(be, np.full(1, -int(k), np.int32), I32)
    leaf, nl = palette_split(be, keys, off, cnt, lab, ncl, mc, max_rows=n)
    check_counts("rhccq_palette_split", nl)
    return leaf.cpu().numpy().astype(np.int64)


def minibatch_labels(be: Backend, colors, quality: float):
    """``MiniBatchKMeans(n_clusters=ceil(n*quality/100/10), batch_size=1000, random_state=42,
    n_init='auto').fit_predict(colors.astype(float))`` — the operator the reference calls at
    clustering.py:207-218 for palettes of 10 000 colours and more — through rhccq_palette_minibatch on one
    palette.  Returns (int64 labels, k)."""
    import numpy as np
    c = np.ascontiguousarray(np.asarray(colors).astype(np.uint8).reshape(-1, 3)).astype(np.int64)
    n = int(c.shape[0])
    if (c == 0).all(axis=1).any():
        raise NotImplementedError("black rows never reach MiniBatchKMeans in the reference (clustering.py:185-192)")
    keys = _as_dev(be, ((c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]).astype(np.int32), I32)
    off = _as_dev(be, np.zeros(1, np.int32), I32)
    cnt = _as_dev(be, np.full(1, n, np.int32), I32)
    lab = be.zeros((n,), I32)
    ncl = _as_dev(be, np.full(1, -4, np.int32), I32)               # the status rhccq_palette_dbscan gives such a palette
    q = _as_dev(be, np.full(1, float(quality), np.float64), torch.float64)
    palette_minibatch(be, keys, off, cnt, q, lab, ncl, max_rows=n)
    check_counts("rhccq_palette_minibatch", ncl)
    return lab.cpu().numpy().astype(np.int64), int(ncl.cpu().numpy()[0])


def palette_finish(be: Backend, pal_keys, pal_off, pal_cnt, leaf, n_leaves, *, max_rows: int):
    """Truncated means — rhccq_palette_finish.  Returns new_keys (same layout as pal_keys)."""
    P = pal_cnt.numel()
    new_keys = be.zeros((pal_keys.numel(),), I32)
    need = be.cdll.rhccq_palette_finish_workspace_bytes(max_rows)
    ws, ws_bytes = be.workspace(need, P)
    be.call("rhccq_palette_finish", be.ptr(pal_keys), be.ptr(pal_off), be.ptr(pal_cnt), P, be.ptr(leaf),
            be.ptr(n_leaves), be.ptr(new_keys), int(max_rows), be.ptr(ws), ws_bytes, be.stream())
    return new_keys


def quality_params(be: Backend, quality, max_rows: int) -> dict:
    """Device-resident radius parameters of a batch of problems from their qualities (host sequence):
    eps = 128 - 1.28 q (0 -> 1, clustering.py:127,131-132), its integer threshold and tie flag, and the
    slot bound of the DBSCAN cell table.  Worth caching: it only depends on the qualities."""
    import numpy as np
    q = np.asarray(quality, dtype=np.float64).reshape(-1)
    uq, inv = np.unique(q, return_inverse=True)
    eps_u = np.array([(128 - 1.28 * float(x)) or 1 for x in uq], dtype=np.float64)
    tt_u = [eps_threshold(float(e)) for e in eps_u]
    max_slots = max([be.cdll.rhccq_palette_dbscan_slots(t[0], int(max_rows)) for t in tt_u] + [1])
    thr = np.array([t[0] for t in tt_u], dtype=np.int32)[inv]
    tie = np.array([t[1] for t in tt_u], dtype=np.int32)[inv]
    return {"q": _as_dev(be, q, torch.float64), "eps": _as_dev(be, eps_u[inv], torch.float64),
            "thr": _as_dev(be, thr, I32), "tie": _as_dev(be, tie, I32), "max_slots": int(max_slots), "n": int(q.size)}


def cluster_palettes(be: Backend, pal_keys, pal_off, pal_cnt, quality, *, max_rows: int, max_cpc=None, params=None):
    """a2 + a3' + a3/a4 + means for a batch of palettes on the device.

    ``quality``: host sequence, one per problem (eps and the slot bound come
    from it on the host; max_cpc from the device-resident counts); or pass
    ``params`` = a cached `quality_params` result.
    Returns dict(labels, n_clusters, leaf, n_leaves, new_keys, max_cpc).
    """
    P = pal_cnt.numel()
    if params is None:
        params = quality_params(be, quality, max_rows)
    labels, ncl = palette_dbscan(be, pal_keys, pal_off, pal_cnt, params["thr"], params["tie"], params["eps"],
                                 max_rows=max_rows, max_slots=params["max_slots"])
    q_dev = params["q"]
    if max_rows >= 10000:                                          # clustering.py:207: the MiniBatchKMeans branch can occur
        palette_minibatch(be, pal_keys, pal_off, pal_cnt, q_dev, labels, ncl, max_rows=max_rows)
    if max_cpc is None:
        max_cpc = cluster_params(be, pal_cnt, q_dev)
    leaf, nl = palette_split(be, pal_keys, pal_off, pal_cnt, labels, ncl, max_cpc, max_rows=max_rows)
    new_keys = palette_finish(be, pal_keys, pal_off, pal_cnt, leaf, nl, max_rows=max_rows)
    return {"labels": labels, "n_clusters": ncl, "leaf": leaf, "n_leaves": nl, "new_keys": new_keys,
            "max_cpc": max_cpc, "P": P}


def excl_scan(be: Backend, counts):
    n = counts.numel()
    out = be.empty((n + 1,), I32)
    be.call("rhccq_excl_scan", be.ptr(counts), n, be.ptr(out), be.stream())
    return out


def remap_first(be: Backend, seg, plane, crops, pal_off, leaf, n_leaves, new_keys, ent_off, n_entries: int,
                *, idx_bytes: int, max_leaves: int):
    """LUT remap of the index plane + entry table (colour, first raster position)."""
    _, B, H, W = plane.shape
    P = crops.shape[0]
    ent_color = be.zeros((max(n_entries, 1),), I32)
    ent_fpos = be.empty((max(n_entries, 1),), I32)
    ent_fpos.fill_(-1)                                              # 0xFFFFFFFF
    be.call("rhccq_remap_first", be.ptr(seg), B, H, W, be.ptr(crops), P, be.ptr(pal_off), be.ptr(leaf),
            be.ptr(n_leaves), be.ptr(new_keys), be.ptr(ent_off), be.ptr(plane), idx_bytes, be.ptr(ent_color),
            be.ptr(ent_fpos), int(max(max_leaves, 1)), be.stream())
    return ent_color, ent_fpos


def merge_level(be: Backend, color_in, fpos_in, comp_start, comp_cnt, grp_comp_off, n_groups: int,
                out_capacity: int, *, max_entries: int, max_comps: int):
    """a5 on entries — rhccq_merge_level."""
    color_out = be.zeros((max(out_capacity, 1),), I32)
    fpos_out = be.empty((max(out_capacity, 1),), I32)
    fpos_out.fill_(-1)
    out_off = be.empty((n_groups + 1,), I32)
    out_cnt = be.zeros((max(n_groups, 1),), I32)
    out_present = be.zeros((max(n_groups, 1),), I32)
    emap = be.zeros((max(color_in.numel(), 1),), I32)
    need = be.cdll.rhccq_merge_level_workspace_bytes(int(max(max_entries, 1)), int(max(max_comps, 1)))
    ws, ws_bytes = be.workspace(need, n_groups)
    be.call("rhccq_merge_level", be.ptr(color_in), be.ptr(fpos_in), be.ptr(comp_start), be.ptr(comp_cnt),
            be.ptr(grp_comp_off), n_groups, be.ptr(color_out), be.ptr(fpos_out), be.ptr(out_off), be.ptr(out_cnt),
            be.ptr(out_present), be.ptr(emap), int(max(max_entries, 1)), int(max(max_comps, 1)), be.ptr(ws),
            ws_bytes, be.stream())
    return {"color": color_out, "fpos": fpos_out, "off": out_off, "cnt": out_cnt, "present": out_present,
            "map": emap}


def first_min(be: Backend, off, cnt, n_leaves, leaf, fpos_in):
    n_groups = cnt.numel()
    fpos_out = be.empty((fpos_in.numel(),), I32)
    fpos_out.fill_(-1)
    be.call("rhccq_first_min", be.ptr(off), be.ptr(cnt), be.ptr(n_leaves), n_groups, be.ptr(leaf),
            be.ptr(fpos_in), be.ptr(fpos_out), be.stream())
    return fpos_out


def compose_final(be: Backend, n_segments, n_leaves1, ent_off0, seg_region, region_group, group_image,
                  A, Bm, leaf2, color2, Cm, leaf3, n_entries: int):
    ent_final = be.empty((max(n_entries, 1),), I32)
    be.call("rhccq_compose_final", n_segments, be.ptr(n_leaves1), be.ptr(ent_off0), be.ptr(seg_region),
            be.ptr(region_group), be.ptr(group_image), be.ptr(A["off"]), be.ptr(A["map"]), be.ptr(Bm["off"]),
            be.ptr(Bm["map"]), be.ptr(leaf2), be.ptr(color2), be.ptr(Cm["off"]), be.ptr(Cm["map"]), be.ptr(leaf3),
            be.ptr(Cm["present"]), be.ptr(ent_final), be.stream())
    return ent_final


def paint(be: Backend, seg, crops, ent_off, ent_final, plane, out_plane, *, cls: int, idx_bytes: int):
    _, B, H, W = plane.shape
    be.call("rhccq_paint", be.ptr(seg), B, H, W, be.ptr(crops), crops.shape[0], be.ptr(ent_off), be.ptr(ent_final),
            int(cls), be.ptr(plane), idx_bytes, be.ptr(out_plane), be.stream())


def check_counts(name: str, counts) -> None:
    """Raise when a kernel flagged a problem (negative counter) — never continue on a refusal."""
    bad = counts[counts < 0]
    if bad.numel():
        code = int(bad[0])
        why = {-1: "capacity bound exceeded", -2: "upstream error / random table too short",
               -3: "more palette rows than the index type can address",
               -4: ">= 10000 non-black colours: the palette needs rhccq_palette_minibatch "
                   "(encoder/compression/clustering.py:207-218) before the split"}.get(code, "unknown")
        raise RhccqError(f"{name}: {bad.numel()} problem(s) refused, first code {code} ({why})")
