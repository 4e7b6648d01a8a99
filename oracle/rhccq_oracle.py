"""numpy restatement of the reference's hierarchical palette quantiser.

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

Every function names the reference lines it restates; paths are relative to
/root/reference.  Dicts carry the reference's keys but
hold numpy arrays where the reference holds Python lists: ``palette`` is
``uint8 [m,3]``, ``indices`` is a flat integer array.  ``to_lists`` converts to
the reference's list form when a test needs to feed the reference itself.

Two third-party calls sit inside the reference path (scikit-learn, unpinned in
requirements.txt:6): ``DBSCAN`` and ``KMeans``.  Both are restated here
(``dbscan_labels`` below, ``oracle/kmeans_restated.py``) and both can be
switched to scikit-learn itself (``dbscan_impl='sklearn'``,
``kmeans_impl='sklearn'``) so that the tests can tell a disagreement of the
restatement from a disagreement of the product.  ``MiniBatchKMeans`` (the
>= 10 000-colour branch, clustering.py:207-218) is restated in
``oracle/minibatch_restated.py`` (``minibatch_impl='sklearn'`` switches to
scikit-learn).
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np

from . import kmeans_restated, minibatch_restated

BLACK_KEY = 0


# --------------------------------------------------------------------------- helpers
def pack_rgb(rgb: np.ndarray) -> np.ndarray:
    """uint8 [...,3] -> uint32 key R<<16|G<<8|B (ascending key == lexicographic RGB)."""
    a = np.asarray(rgb).astype(np.uint32)
    return (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]


def unpack_rgb(key: np.ndarray) -> np.ndarray:
    k = np.asarray(key).astype(np.uint32)
    return np.stack([(k >> 16) & 255, (k >> 8) & 255, k & 255], axis=-1).astype(np.uint8)


def to_lists(comp: dict) -> dict:
    """Reference form of a component dict: palette list-of-lists, indices list-of-int."""
    out = dict(comp)
    out["palette"] = np.asarray(comp["palette"]).reshape(-1, 3).astype(int).tolist()
    out["indices"] = np.asarray(comp["indices"]).astype(int).ravel().tolist()
    return out


def from_reference(comp: dict) -> dict:
    """Array form of a dict produced by the reference (lists or tuples inside)."""
    out = dict(comp)
    out["palette"] = np.asarray(comp["palette"], dtype=np.uint8).reshape(-1, 3)
    out["indices"] = np.asarray(comp["indices"], dtype=np.int64).ravel()
    return out


# --------------------------------------------------------------------------- a1
def get_all_unique_colors(region_image: np.ndarray, top_left_coords) -> dict | None:
    """encoder/compression/clustering.py:4-103.

    Palette = ``np.unique(pixels, axis=0)`` (lexicographic R,G,B, :21-23);
    index per pixel = row of its colour (:31-48).
    """
    if region_image is None or region_image.size == 0:       # :9-10
        return None
    h, w, _ = region_image.shape
    keys = pack_rgb(region_image.reshape(-1, 3))
    uniq, inv = np.unique(keys, return_inverse=True)
    n = int(uniq.size)
    total = h * w
    bytes_per_index = 1 if n <= 256 else 2                   # :57-62
    compressed = n * 3 + total * bytes_per_index + 50        # :65-69
    return {
        "method": "exact_colors",
        "top_left": top_left_coords,
        "shape": (h, w),
        "palette": unpack_rgb(uniq),
        "indices": inv.astype(np.int64).ravel(),
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


# --------------------------------------------------------------------------- a2
def compute_clustering_params(n_colors, quality, color_space="rgb"):
    """encoder/compression/clustering.py:108-135 (color_space is ignored there too)."""
    eps = 128 - 1.28 * quality                                                   # :127
    max_cpc = math.ceil((-(quality / 100) * n_colors + n_colors) / quality)      # :129
    if eps == 0:
        eps = 1
    if max_cpc == 0:
        max_cpc = 1
    return eps, 1, max_cpc


# --------------------------------------------------------------------------- a3'
def eps_threshold(eps: float) -> tuple[int, bool]:
    """Integer form of sklearn's radius predicate on 0..255 colours.

    The reference calls ``DBSCAN(eps=eps/255.0).fit_predict(palette/255.0)``
    (clustering.py:205,233-235); the KD-tree accepts a pair when
    ``sum(((a-b)/255)^2) <= (eps/255)^2`` in float64
    (sklearn/neighbors/_binary_tree.pxi.tp:1952-1957).  For integer colours the
    left side times 255^2 is the integer d2, so the test is ``d2 <= floor(eps^2)``
    unless ``eps^2`` is itself an integer, where a pair with ``d2 == eps^2`` is a
    floating-point tie (returned flag) that must be evaluated as sklearn does.
    """
    e2 = float(eps) * float(eps)
    thr = int(math.floor(e2 + 1e-9))
    tie = abs(e2 - round(e2)) < 1e-9
    if tie:
        thr = int(round(e2))
    return thr, tie


def _sk_pair_accept(a: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    """float64 evaluation of the radius predicate exactly as the KD-tree leaf does."""
    an = a.astype(np.float64) / 255.0
    bn = b.astype(np.float64) / 255.0
    d = an - bn
    s = d[..., 0] * d[..., 0]
    s = s + d[..., 1] * d[..., 1]
    s = s + d[..., 2] * d[..., 2]
    r = float(eps) / 255.0
    return s <= r * r


def _components_min_index(adj_rows: Callable[[int, int], np.ndarray], n: int, step: int) -> np.ndarray:
    """Connected components of a graph given row blocks of its adjacency; root = min index."""
    lab = np.arange(n, dtype=np.int64)
    changed = True
    while changed:
        changed = False
        new = lab.copy()
        for lo in range(0, n, step):
            a = adj_rows(lo, min(n, lo + step))                  # bool [rows, n]
            cand = np.where(a, lab[None, :], n).min(axis=1)
            new[lo:lo + a.shape[0]] = np.minimum(new[lo:lo + a.shape[0]], cand)
        new = new[new]                                           # pointer jumping
        if not np.array_equal(new, lab):
            changed = True
            lab = new
    return lab


def dbscan_labels(points: np.ndarray, eps: float, min_samples: int = 1,
                  colour_scale: bool = False) -> np.ndarray:
    """Brute-force restatement of ``sklearn.cluster.DBSCAN(...).fit_predict``.

    sklearn/cluster/_dbscan.py:397-470 and _dbscan_inner.pyx: the neighbourhood
    of a point contains the point itself; core <=> |neighbourhood| >= min_samples;
    clusters are numbered by their lowest *core* index (the DFS seeds ascend);
    a non-core point reached from several clusters keeps the lowest-numbered
    one (it is labelled by the first cluster expanded); the rest is noise (-1).

    ``colour_scale=True``: ``points`` are uint8 colours and ``eps`` is on the
    0..255 scale, predicate as in ``eps_threshold`` (the reference's call).
    Otherwise points are float32/float64 [n,D] and the predicate is the
    KD-tree's: float64 ``sum_d (x_d - y_d)^2 <= eps^2`` summed in dimension
    order.  O(n^2) time, row-blocked memory: for tests up to ~20 000 points.
    """
    p = np.asarray(points)
    n = p.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    step = max(1, (1 << 24) // n)
    if colour_scale:
        pi = p.astype(np.int64).reshape(n, 3)
        thr, tie = eps_threshold(eps)

        def adj_rows(lo, hi):
            d = pi[lo:hi, None, :] - pi[None, :, :]
            d2 = (d * d).sum(axis=2)
            a = d2 < thr if tie else d2 <= thr
            if tie:
                eq = d2 == thr
                if eq.any():
                    r, c = np.nonzero(eq)
                    a[r, c] = _sk_pair_accept(pi[lo + r], pi[c], eps)
            return a
    else:
        pf = p.astype(np.float64)
        r2 = float(eps) * float(eps)

        def adj_rows(lo, hi):
            s = np.zeros((hi - lo, n), dtype=np.float64)
            for d in range(pf.shape[1]):
                t = pf[lo:hi, None, d] - pf[None, :, d]
                s = s + t * t
            return s <= r2

    if min_samples <= 1:
        # every point is core (its neighbourhood holds itself): labels are the connected
        # components of the eps-graph, numbered by lowest index.  Flood fill from each seed
        # against the still unlabelled points only.
        labels = np.full(n, -1, dtype=np.int64)
        todo = np.arange(n)
        c = 0
        while todo.size:
            seed = todo[0]
            labels[seed] = c
            todo = todo[1:]
            frontier = [seed]
            while frontier and todo.size:
                i = frontier.pop()
                hit = adj_rows(i, i + 1)[0][todo]
                if hit.any():
                    got = todo[hit]
                    labels[got] = c
                    frontier.extend(got.tolist())
                    todo = todo[~hit]
            c += 1
        return labels

    counts = np.zeros(n, dtype=np.int64)
    for lo in range(0, n, step):
        counts[lo:lo + step] = adj_rows(lo, min(n, lo + step)).sum(axis=1)
    core = counts >= min_samples

    def core_adj(lo, hi):
        a = adj_rows(lo, hi) & core[None, :] & core[lo:hi, None]
        idx = np.arange(lo, hi)
        a[idx - lo, idx] = True
        return a

    root = _components_min_index(core_adj, n, step)
    labels = np.full(n, -1, dtype=np.int64)
    core_roots = np.unique(root[core])                           # ascending lowest core index
    rank = {int(r): i for i, r in enumerate(core_roots)}
    if core.any():
        labels[core] = np.array([rank[int(r)] for r in root[core]], dtype=np.int64)
    if not core.all():
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            nc = np.flatnonzero(~core[lo:hi])
            if nc.size == 0:
                continue
            a = adj_rows(lo, hi)[nc] & core[None, :]
            cand = np.where(a, labels[None, :], n + 1).min(axis=1)
            labels[lo + nc] = np.where(cand > n, -1, cand)
    return labels


def _dbscan_sklearn(colors_u8: np.ndarray, eps: float, min_samples: int) -> np.ndarray:
    from sklearn.cluster import DBSCAN
    x = colors_u8.astype(float) / 255.0                                          # clustering.py:205
    return DBSCAN(eps=eps / 255.0, min_samples=min_samples, metric="euclidean").fit_predict(x)


# --------------------------------------------------------------------------- a4
def _kmeans_sklearn(colors_u8: np.ndarray, k: int) -> np.ndarray:
    from sklearn.cluster import KMeans
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return KMeans(n_clusters=k, random_state=42, n_init="auto").fit_predict(colors_u8.astype(float))


def split_by_luminance(ids: np.ndarray, colors: np.ndarray, max_cpc: int) -> list[np.ndarray]:
    """encoder/compression/clustering.py:778-801 (K-Means failure fallback)."""
    n = len(ids)
    if n <= max_cpc:
        return [ids]
    c = colors[ids].astype(np.float64)
    lum = 0.299 * c[:, 0] + 0.587 * c[:, 1] + 0.114 * c[:, 2]
    order = np.argsort(lum)
    n_splits = max(2, (n + max_cpc - 1) // max_cpc)
    return [s for s in np.array_split(ids[order], n_splits) if len(s) > 0]


def split_large_cluster(ids: np.ndarray, colors: np.ndarray, max_cpc: int,
                        kmeans: Callable[[np.ndarray, int], np.ndarray]) -> list[np.ndarray]:
    """encoder/compression/clustering.py:720-775, on palette-row ids instead of colours.

    ``ids`` ascend (the reference builds ``cluster_colors = palette[ascending rows]``);
    each returned split keeps that order, and the splits come back in the
    reference's order: K-Means label order with recursive splits expanded in place.
    """
    n = len(ids)
    if n <= max_cpc:                                             # :735
        return [ids]
    n_splits = max(2, (n + max_cpc - 1) // max_cpc)              # :739
    n_splits = min(n_splits, n)                                  # :742
    if n <= 2 or n_splits < 2:                                   # :745
        return [ids]
    labels = np.asarray(kmeans(colors[ids], n_splits))
    out: list[np.ndarray] = []
    for i in range(n_splits):                                    # :755-758
        part = ids[labels == i]
        if part.size == 0:
            continue
        if part.size > max_cpc and part.size < n:                # :763-767
            out.extend(split_large_cluster(part, colors, max_cpc, kmeans))
        else:
            # part.size == n would recurse forever in the reference (RecursionError
            # -> worker returns None -> whole cluster averaged, :335-343); one leaf
            # holding the whole cluster is the same result.
            out.append(part)
    return out


# --------------------------------------------------------------------------- a3
def cluster_palette_colors_parallel(quality, compressed_data: dict, eps=10.0, min_samples=2,
                                    max_colors_per_cluster=5, num_workers=None, *,
                                    dbscan_impl: str = "restated",
                                    kmeans_impl: str | Callable = "restated",
                                    minibatch_impl: str | Callable = "restated") -> dict:
    """encoder/compression/clustering.py:160-437.

    Large clusters are consumed in submission order (ascending DBSCAN label);
    the reference consumes them in thread-completion order (:458), which is the
    one nondeterministic step of the path and is normalised away on both sides.
    """
    palette = np.asarray(compressed_data["palette"], dtype=np.uint8).reshape(-1, 3)     # :171
    indices = np.asarray(compressed_data["indices"]).astype(np.int64).ravel()            # :172
    h, w = compressed_data["shape"]
    n_orig = len(palette)
    keys = pack_rgb(palette)
    black = np.flatnonzero(keys == BLACK_KEY)                                            # :185-192
    non_black = np.flatnonzero(keys != BLACK_KEY)
    if non_black.size == 0:                                                              # :197-199
        return compressed_data
    nb_pal = palette[non_black]
    if non_black.size >= 10000:                                                          # :207-218
        n_clusters = math.ceil(len(nb_pal) * (quality / 100) / 10)
        if minibatch_impl == "sklearn":
            from sklearn.cluster import MiniBatchKMeans
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                labels = MiniBatchKMeans(n_clusters=n_clusters, batch_size=1000, random_state=42,
                                         n_init="auto").fit_predict(nb_pal.astype(float))
        elif callable(minibatch_impl):
            labels = minibatch_impl(nb_pal, n_clusters)
        else:
            labels = minibatch_restated.minibatch_labels(nb_pal, n_clusters)
    elif dbscan_impl == "sklearn":
        labels = _dbscan_sklearn(nb_pal, eps, min_samples)                               # :233-235
    else:
        labels = dbscan_labels(nb_pal, eps, min_samples, colour_scale=True)
    labels = np.asarray(labels).astype(np.int64)

    if callable(kmeans_impl):
        kmeans = kmeans_impl
    elif kmeans_impl == "sklearn":
        kmeans = _kmeans_sklearn
    else:
        kmeans = kmeans_restated.kmeans_labels

    new_palette: list[np.ndarray] = []
    lut = np.zeros(n_orig, dtype=np.int64)                                               # :373 (zeros)
    for b in black:                                                                      # :253-255
        lut[b] = len(new_palette)
        new_palette.append(palette[b])
    for rel in np.flatnonzero(labels == -1):                                             # :258-264
        lut[non_black[rel]] = len(new_palette)
        new_palette.append(palette[non_black[rel]])
    uniq = np.unique(labels[labels >= 0])                                                # :273-275 ascending
    large = []
    for lab in uniq:
        ids = non_black[labels == lab]                                                   # ascending rows
        if ids.size > max_colors_per_cluster:                                            # :284
            large.append(ids)
        else:                                                                            # :304-310
            lut[ids] = len(new_palette)
            new_palette.append((palette[ids].astype(np.int64).sum(axis=0) // ids.size).astype(np.uint8))
    first_row: dict[int, int] = {}
    for i, k in enumerate(keys.tolist()):
        first_row.setdefault(k, i)
    for ids in large:                                                                    # :330-355
        for part in split_large_cluster(ids, palette, max_colors_per_cluster, kmeans):
            new_idx = len(new_palette)
            new_palette.append((palette[part].astype(np.int64).sum(axis=0) // part.size).astype(np.uint8))
            for i in part:                                                               # find_color_index :803-808
                lut[first_row[int(keys[i])]] = new_idx
    new_pal = np.asarray(new_palette, dtype=np.uint8).reshape(-1, 3)
    m = len(new_pal)
    new_indices = lut.astype(np.uint16)[indices].astype(np.int64)                        # :373-377 (uint16 table)
    total = h * w
    original_size = compressed_data.get("original_size", total * 3)
    bpi = 1 if m <= 256 else 2
    new_size = m * 3 + total * bpi + 100
    return {
        "method": "clustered_colors",
        "top_left": compressed_data["top_left"],
        "shape": (h, w),
        "palette": new_pal,
        "indices": new_indices,
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


# --------------------------------------------------------------------------- a5
def merge_region_components_simple(region_components: Sequence[dict], roi_bbox) -> list[dict]:
    """encoder/compression/merging.py:8-120.

    Components are painted in reversed list order (:52), black never paints
    (:76), a colour enters the palette when the paint sequence first meets it
    (:77-79) even if that pixel is overwritten later.
    """
    if not region_components:                                    # :13-14
        return []
    if len(region_components) == 1:                              # :16-21
        single = dict(region_components[0])
        if "actual_colors" not in single:
            single["actual_colors"] = len(single.get("palette", []))
        return [single]
    minr, minc, maxr, maxc = roi_bbox
    H, W = maxr - minr, maxc - minc
    canvas = np.zeros((H, W), dtype=np.uint32)                   # :35
    pal_keys = [BLACK_KEY]                                       # :42-44
    index_of = {BLACK_KEY: 0}
    for seg in reversed(region_components):                      # :52
        sh, sw = seg["shape"]
        pal = pack_rgb(np.asarray(seg["palette"], dtype=np.uint8).reshape(-1, 3))
        idx = np.asarray(seg["indices"]).astype(np.int64).reshape(sh, sw)
        if len(pal) == 0:                                        # :69-72: no index is below len(palette)
            continue
        r0 = seg["top_left"][0] - minr
        c0 = seg["top_left"][1] - minc
        rr, cc = np.meshgrid(np.arange(sh) + r0, np.arange(sw) + c0, indexing="ij")
        ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W) & (idx < len(pal))   # :69-72
        col = np.where(ok, pal[np.minimum(idx, len(pal) - 1)], BLACK_KEY)
        ok &= col != BLACK_KEY                                   # :76
        flat = col[ok]                                           # raster order
        if flat.size == 0:
            continue
        u, first = np.unique(flat, return_index=True)
        for k in u[np.argsort(first, kind="stable")].tolist():   # :77-79
            if k not in index_of:
                index_of[k] = len(pal_keys)
                pal_keys.append(k)
        lut = np.array([index_of[k] for k in u.tolist()], dtype=np.uint32)
        canvas[rr[ok], cc[ok]] = lut[np.searchsorted(u, flat)]   # :81
    n = len(pal_keys)
    dt = np.uint8 if n <= 256 else (np.uint16 if n <= 65536 else np.uint32)   # :99-104
    return [{
        "top_left": (minr, minc),
        "shape": (H, W),
        "palette": unpack_rgb(np.asarray(pal_keys, dtype=np.uint32)),
        "indices": canvas.astype(np.int64).ravel(),
        "indices_dtype": str(dt),
        "method": "merged",
        "actual_colors": n,
        "encoding": "roi_merged",
    }]


# --------------------------------------------------------------------------- a6
def segment_component(image_rgb: np.ndarray, region: dict, segments: np.ndarray, segment_id: int,
                      quality, **cluster_kw) -> dict | None:
    """One pass of the per-segment loop, encoder/compression/subregions.py:315-449."""
    minr, minc, maxr, maxc = region["bbox"]
    region_image = image_rgb[minr:maxr, minc:maxc]
    mask = (segments == segment_id) & region["bbox_mask"]        # :317
    rows, cols = np.where(mask)                                  # :340
    if rows.size == 0:                                           # :342-343
        return None
    pad = 2                                                      # :350-355
    h, w = region_image.shape[:2]
    r0, r1 = max(0, rows.min() - pad), min(h - 1, rows.max() + pad)
    c0, c1 = max(0, cols.min() - pad), min(w - 1, cols.max() + pad)
    crop = region_image[r0:r1 + 1, c0:c1 + 1]
    cmask = mask[r0:r1 + 1, c0:c1 + 1]
    seg_img = np.zeros_like(crop)                                # :371-372
    seg_img[cmask] = crop[cmask]
    px = crop[cmask].copy()                                      # :391
    keys = pack_rgb(px)
    is_black = keys == BLACK_KEY
    if is_black.any() and (~is_black).any():                     # :395-421
        nb = px[~is_black].astype(np.float64)
        # distance to the black pixel itself == norm of the colour; first minimum wins
        best = int(np.argmin(np.sqrt((nb * nb).sum(axis=1))))
        px[is_black] = px[~is_black][best]
        seg_img[cmask] = px
    comp = get_all_unique_colors(seg_img, (int(r0 + minr), int(c0 + minc)))             # :426
    eps, _, max_cpc = compute_clustering_params(comp["actual_colors"], quality, "lab")  # :436
    return cluster_palette_colors_parallel(quality, comp, eps=eps, min_samples=1,
                                           max_colors_per_cluster=max_cpc, **cluster_kw)  # :443


def subregion_quantization(image_rgb: np.ndarray, subregions: Sequence[dict], quality=10,
                           subregion_type=None, debug=False, **cluster_kw) -> list:
    """encoder/compression/subregions.py:90-683 with the SLIC label map supplied.

    Each region dict carries ``bbox``, ``bbox_mask`` (encoder/ROI/roi.py:349-358)
    and ``segments``: the int32 label map ``enhanced_slic_with_texture`` would
    return for it (subregions.py:160), 0 = outside.  Segment order is ascending
    id (encoder/subregions/slic.py:158-162).
    """
    out = []
    for region in subregions:                                    # :98
        comps = []
        segs = region["segments"]
        ids = np.unique(segs)
        for sid in ids[ids != 0]:
            c = segment_component(image_rgb, region, segs, int(sid), quality, **cluster_kw)
            if c is not None:
                comps.append(c)                                  # :634
        if len(comps) > 1:                                       # :639-650
            out.append(merge_region_components_simple(comps, tuple(region["bbox"])))
        else:                                                    # :679
            out.append(comps)
    return out


# --------------------------------------------------------------------------- a7
def _flatten(regions_components) -> list[dict]:
    flat = []                                                    # regions.py:18-29
    for r in regions_components:
        if isinstance(r, dict):
            flat.append(r)
        elif isinstance(r, list):
            flat.extend(x for x in r if isinstance(x, dict))
    return flat


def region_quantization(regions_components, original_image_height, original_image_width,
                        quality=50, **cluster_kw) -> list[dict]:
    """encoder/compression/regions.py:9-70."""
    merged = merge_region_components_simple(
        _flatten(regions_components), (0, 0, original_image_height, original_image_width))[0]
    eps, _, max_cpc = compute_clustering_params(merged["actual_colors"], quality, "lab")
    return [cluster_palette_colors_parallel(quality, merged, eps=eps, min_samples=1,
                                            max_colors_per_cluster=max_cpc, **cluster_kw)]


def optimize_compressed_dtype(comp: dict) -> dict:
    """encoder/compression/compression.py:326-413 (values unchanged; bookkeeping keys only)."""
    if "indices" not in comp:
        return comp
    idx = np.asarray(comp["indices"])
    mx = int(idx.max()) if idx.size else 0
    name = "uint8" if mx < 256 else ("uint16" if mx < 65536 else "uint32")
    out = dict(comp)
    out["indices_dtype"] = name
    out["indices_optimized"] = True
    out["actual_colors"] = len(comp["palette"])
    return out


def quantize_image(image_components, original_image_height, original_image_width,
                   quality=100, **cluster_kw) -> dict:
    """encoder/compression/image.py:243-289."""
    merged = merge_region_components_simple(
        list(image_components), (0, 0, original_image_height, original_image_width))[0]
    eps, _, max_cpc = compute_clustering_params(merged["actual_colors"], quality, "lab")
    comp = cluster_palette_colors_parallel(quality, merged, eps=eps, min_samples=1,
                                           max_colors_per_cluster=max_cpc, **cluster_kw)
    return optimize_compressed_dtype(comp)


def encode_image(image_rgb: np.ndarray, roi_regions, nonroi_regions, roi_quality=20,
                 nonroi_quality=10, **cluster_kw) -> dict:
    """The three-stage schedule of encoder/compression/test.py:100-142."""
    H, W, _ = image_rgb.shape
    s1_roi = subregion_quantization(image_rgb, roi_regions, roi_quality, "ROI", **cluster_kw)
    s1_non = subregion_quantization(image_rgb, nonroi_regions, nonroi_quality, "nonROI", **cluster_kw)
    q2r, q2n = min(100, roi_quality * 2), min(100, nonroi_quality * 2)              # test.py:116-120
    try:                                                                             # :124-128
        roi = region_quantization(s1_roi, H, W, q2r, **cluster_kw)
    except Exception:
        roi = []
    try:
        non = region_quantization(s1_non, H, W, q2n, **cluster_kw)
    except Exception:
        non = []
    return quantize_image(roi + non, H, W, min(100, q2r + q2n), **cluster_kw)       # :139-142


# --------------------------------------------------------------------------- decode / metrics
def decode(comp: dict) -> np.ndarray:
    """decoder/uncompression/uncompression.py:156-218 (the gather at :209)."""
    h, w = comp["shape"]
    pal = np.asarray(comp["palette"], dtype=np.uint8).reshape(-1, 3)
    return pal[np.asarray(comp["indices"]).astype(np.int64).reshape(h, w)]


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """decoder/uncompression/comparison.py:43-44."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else float(10 * np.log10(255.0 * 255.0 / mse))
