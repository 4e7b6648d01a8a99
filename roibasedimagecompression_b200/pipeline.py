"""Device-resident three-stage encode of a batch of images.

This is the fused form of the reference's driver
(/root/reference/encoder/compression/test.py:100-142):

    subregion_quantization(ROI, q) ; subregion_quantization(non-ROI, q')     stage 1
    region_quantization(., 2q) ; region_quantization(., 2q')                 stage 2
    quantize_image(ROI + non-ROI, min(100, 2q + 2q'))                        stage 3

with everything between "RGB + label maps on the device" and "final palette +
index plane on the device" kept in HBM.  Pixels are touched three times
(unique / remap / paint); the three palette clusterings and the three merge
levels (segments -> region -> class canvas -> image) run on per-segment
palette *entries*.  One host synchronisation (the per-segment palette sizes
after stage 1) sizes the entry tables.

Host-side inputs (`SegmentTable`) come from the reference's region dicts
(`table_from_regions`) or, for the synthetic tile segmentation of the
benchmarks, in closed form (`table_from_tiles`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops
from ._lib import Backend, RhccqError

PAD = 2          # encoder/compression/subregions.py:350


@dataclass
class SegmentTable:
    """Segments of a batch, ordered by (image, class, region, segment).

    crops        int32 [P,8]  (image, row0, col0, h, w, segment id, class, 0) — padded tight boxes
    seg_region   int32 [P]    global region index of every segment
    region_group int32 [R]    group = image * K + class of every region
    region_bbox  int32 [R,4]  (minr, minc, maxr, maxc)
    n_valid      int64 [P]    pixels of every segment
    """
    B: int
    K: int
    H: int
    W: int
    crops: np.ndarray
    seg_region: np.ndarray
    region_group: np.ndarray
    region_bbox: np.ndarray
    n_valid: np.ndarray
    qualities: tuple = (20.0, 10.0)      # stage-1 quality per class (ROI, non-ROI)

    def dev(self, be, name: str, make):
        """Device copy of a host-side constant of this table, made once per backend."""
        cache = self.__dict__.setdefault("_dev_cache", {})
        key = (id(be), name)
        if key not in cache:
            cache[key] = make()
        return cache[key]

    @property
    def P(self) -> int:
        return int(self.crops.shape[0])

    @property
    def R(self) -> int:
        return int(self.region_group.shape[0])


def table_from_regions(image_shape, regions_per_class, qualities, B: int = 1, image_index: int = 0,
                       tables=None) -> tuple[SegmentTable, np.ndarray]:
    """Segment table + label maps [K,1,H,W] from the reference's region dicts.

    ``regions_per_class``: list over classes (ROI first) of lists of region dicts with
    ``bbox`` (minr, minc, maxr, maxc), ``bbox_mask`` (encoder/ROI/roi.py:349-358) and
    ``segments``, the int32 label map of the bbox that ``enhanced_slic_with_texture``
    returns (encoder/compression/subregions.py:160; 0 = outside).  Box rules:
    subregions.py:317-359 (mask, tight box, pad 2, clamp to the region's bbox).
    """
    H, W = image_shape[:2]
    K = len(regions_per_class)
    label = np.zeros((K, 1, H, W), dtype=np.int32)
    crops, seg_region, region_group, region_bbox, n_valid = [], [], [], [], []
    next_id = 1
    for k, regions in enumerate(regions_per_class):
        for region in regions:
            minr, minc, maxr, maxc = (int(v) for v in region["bbox"])
            segs = np.asarray(region["segments"])
            mask = np.asarray(region["bbox_mask"]).astype(bool)
            r_idx = len(region_group)
            region_group.append(image_index * K + k)
            region_bbox.append((minr, minc, maxr, maxc))
            h, w = maxr - minr, maxc - minc
            ids = np.unique(segs)
            for sid in ids[ids != 0]:                              # ascending id: slic.py:158-162
                m = (segs == sid) & mask                           # subregions.py:317
                rows, cols = np.where(m)
                if rows.size == 0:                                 # :342-343
                    continue
                r0, r1 = max(0, rows.min() - PAD), min(h - 1, rows.max() + PAD)
                c0, c1 = max(0, cols.min() - PAD), min(w - 1, cols.max() + PAD)
                view = label[k, 0, minr:maxr, minc:maxc]
                if np.any(view[m] != 0):
                    raise ValueError("segments of one class overlap; masks of a class must be disjoint")
                view[m] = next_id
                crops.append((image_index, minr + r0, minc + c0, r1 - r0 + 1, c1 - c0 + 1, next_id, k, 0))
                seg_region.append(r_idx)
                n_valid.append(rows.size)
                next_id += 1
    tab = SegmentTable(B=B, K=K, H=H, W=W,
                       crops=np.asarray(crops, dtype=np.int32).reshape(-1, 8),
                       seg_region=np.asarray(seg_region, dtype=np.int32),
                       region_group=np.asarray(region_group, dtype=np.int32),
                       region_bbox=np.asarray(region_bbox, dtype=np.int32).reshape(-1, 4),
                       n_valid=np.asarray(n_valid, dtype=np.int64),
                       qualities=tuple(float(q) for q in qualities))
    return tab, label


def table_from_tiles(B: int, H: int, W: int, tile: int = 64, qualities=(20.0, 10.0)) -> tuple[SegmentTable, np.ndarray]:
    """The synthetic segmentation of SURVEY.md 8d in closed form: square tiles, checkerboard classes,
    one region per class and image (its bbox = the extent of the class's tiles).  Returns the table
    and the label maps [2,1,H,W] of ONE image (identical for every image of the batch)."""
    from .synth import tile_regions
    roi, non = tile_regions(H, W, tile)
    t1, lab = table_from_regions((H, W), [roi, non], qualities)
    P1, R1 = t1.P, t1.R
    crops = np.tile(t1.crops, (B, 1))
    crops[:, 0] = np.repeat(np.arange(B, dtype=np.int32), P1)
    # order by (image, class, region, segment): table_from_regions emits class-major within an image
    seg_region = np.concatenate([t1.seg_region + b * R1 for b in range(B)]).astype(np.int32)
    region_group = np.concatenate([t1.region_group + b * t1.K for b in range(B)]).astype(np.int32)
    tab = SegmentTable(B=B, K=t1.K, H=H, W=W, crops=crops, seg_region=seg_region, region_group=region_group,
                       region_bbox=np.tile(t1.region_bbox, (B, 1)), n_valid=np.tile(t1.n_valid, B),
                       qualities=t1.qualities)
    return tab, lab


@dataclass
class EncodeResult:
    """Device-resident result of `encode_batch`."""
    indices: torch.Tensor            # uint16-valued int16 [B,H,W]
    palette_keys: torch.Tensor       # int32 [capacity] packed RGB keys
    palette_off: torch.Tensor        # int32 [B+1]
    palette_cnt: torch.Tensor        # int32 [B]
    stage: dict = field(default_factory=dict)

    def palette(self, b: int) -> np.ndarray:
        o, n = int(self.palette_off[b]), int(self.palette_cnt[b])
        k = self.palette_keys[o:o + n].cpu().numpy().astype(np.uint32)
        return np.stack([(k >> 16) & 255, (k >> 8) & 255, k & 255], axis=-1).astype(np.uint8)

    def index_image(self, b: int) -> np.ndarray:
        return self.indices[b].cpu().numpy().view(np.uint16)


def _dev(be: Backend, a, dtype=torch.int32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).to(be.device)


def _refuse_duplicate_rows(keys, off: int, cnt: int, what: str) -> None:
    """A palette that reaches the next clustering unmerged (merging.py:16-21 copies a single component) may
    hold the same colour in two rows, when two clusters truncate to the same mean.  The reference then maps
    only the first row of a colour (find_color_index, clustering.py:803-808) and leaves the other at index 0;
    the entry tables of this pipeline hold one entry per colour, so that case is refused, never approximated."""
    k = keys[off:off + cnt].cpu().numpy()
    if np.unique(k).size != k.size:
        raise RhccqError(f"{what}: the palette passed on unmerged holds a colour twice "
                         "(reference behaviour: clustering.py:803-808 maps the first row only); not supported")


def _single_component_groups(table: SegmentTable):
    """Groups (image, class) whose only region has a single segment: [(group, segment index)]."""
    seg_per_region = np.bincount(table.seg_region, minlength=table.R)
    reg_per_group = np.bincount(table.region_group, minlength=table.B * table.K)
    out = []
    for r in np.flatnonzero(seg_per_region == 1):
        g = int(table.region_group[r])
        if reg_per_group[g] == 1:
            out.append((g, int(np.flatnonzero(table.seg_region == r)[0])))
    return out


def stage1(be: Backend, images, labels, table: SegmentTable) -> dict:
    """Stage 1 for every segment of the batch and the per-region merge
    (encoder/compression/subregions.py:315-449, :634-683)."""
    B, K, H, W, P, R = table.B, table.K, table.H, table.W, table.P, table.R
    if tuple(images.shape) != (B, H, W, 3) or images.dtype != torch.uint8:
        raise ValueError(f"images must be uint8 [{B},{H},{W},3]")
    if tuple(labels.shape) != (K, B, H, W) or labels.dtype != torch.int32:
        raise ValueError(f"labels must be int32 [{K},{B},{H},{W}]")
    if P == 0:
        raise IndexError("no segments: the reference fails here too (regions.py:39 indexes an empty list)")
    q1 = np.asarray(table.qualities, dtype=np.float64)
    crops_h = table.crops
    cls_h = crops_h[:, 6]
    cap = (crops_h[:, 3].astype(np.int64) * crops_h[:, 4])
    pal_off_h = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(cap, out=pal_off_h[1:])
    if pal_off_h[-1] >= 2 ** 31:
        raise RhccqError("batch too large for int32 palette offsets; split the batch")
    max_valid = int(table.n_valid.max())
    idxb = 2 if max_valid + 1 <= 65535 else 4                      # a segment's palette has at most n_valid + 1 rows
    crops = table.dev(be, "crops", lambda: _dev(be, crops_h))
    pal_off = table.dev(be, "pal_off", lambda: _dev(be, pal_off_h[:-1]))
    pal_keys, pal_cnt, plane = ops.unique_index(be, images, labels, crops, pal_off, int(pal_off_h[-1]),
                                                idx_bytes=idxb, repaint_black=True, max_valid=max_valid,
                                                n_classes=K)
    s1 = ops.cluster_palettes(be, pal_keys, pal_off, pal_cnt, None, max_rows=max_valid + 1,
                              params=table.dev(be, "q1", lambda: ops.quality_params(be, q1[cls_h], max_valid + 1)))
    nl1_h = s1["n_leaves"][:P].cpu().numpy()                        # the one host synchronisation
    ops.check_counts("stage 1", torch.from_numpy(nl1_h))
    for _, p in table.__dict__.setdefault("_single_groups", _single_component_groups(table)):
        _refuse_duplicate_rows(s1["new_keys"], int(pal_off_h[p]), int(nl1_h[p]), f"stage 1, segment {p}")
    ent_off_h = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(nl1_h, out=ent_off_h[1:])
    E = int(ent_off_h[-1])
    ent_off0 = _dev(be, ent_off_h)
    ent_color0, ent_fpos0 = ops.remap_first(be, labels, plane, crops, pal_off, s1["leaf"], s1["n_leaves"],
                                            s1["new_keys"], ent_off0, E, idx_bytes=idxb,
                                            max_leaves=int(nl1_h.max()))
    # merge level A: segments -> region (subregions.py:639-650)
    reg_first = np.searchsorted(table.seg_region, np.arange(R + 1)).astype(np.int32)   # segments of region r
    ent_per_region = ent_off_h[reg_first[1:]] - ent_off_h[reg_first[:-1]]
    A = ops.merge_level(be, ent_color0, ent_fpos0, ent_off0, s1["n_leaves"],
                        table.dev(be, "reg_first", lambda: _dev(be, reg_first)), R, E + R + 1,
                        max_entries=int(max(ent_per_region.max(), 1)),
                        max_comps=int(max(np.diff(reg_first).max(), 1)))
    return {"pal_keys": pal_keys, "pal_off": pal_off, "pal_cnt": pal_cnt, "plane": plane, "s1": s1, "nl1": nl1_h,
            "ent_off_h": ent_off_h, "ent_off0": ent_off0, "ent_color0": ent_color0, "ent_fpos0": ent_fpos0,
            "A": A, "crops": crops, "reg_first": reg_first, "ent_per_region": ent_per_region, "E": E, "idx_bytes": idxb}


def encode_batch(be: Backend, images, labels, table: SegmentTable, *, keep_stages: bool = False) -> EncodeResult:
    """Three-stage encode of `images` (uint8 [B,H,W,3]) with label maps `labels` (int32 [K,B,H,W]),
    both already on the backend's device."""
    B, K, H, W, P, R = table.B, table.K, table.H, table.W, table.P, table.R
    G = B * K
    st = stage1(be, images, labels, table)
    q1 = np.asarray(table.qualities, dtype=np.float64)
    q2 = np.minimum(100.0, 2.0 * q1)                               # test.py:116-120
    q3 = float(min(100.0, q2.sum()))                               # test.py:139-140
    A, s1, E, ent_off0, crops, plane = st["A"], st["s1"], st["E"], st["ent_off0"], st["crops"], st["plane"]
    ent_per_region = st["ent_per_region"]
    # ---- merge level B: regions -> class canvas (regions.py:18-39)
    region_group_h = table.region_group
    grp_first = np.searchsorted(region_group_h, np.arange(G + 1)).astype(np.int32)   # regions of group g
    capA_cum = np.concatenate([[0], np.cumsum(ent_per_region + 1)])
    ent_per_group = capA_cum[grp_first[1:]] - capA_cum[grp_first[:-1]]
    Bm = ops.merge_level(be, A["color"], A["fpos"], A["off"], A["cnt"],
                         table.dev(be, "grp_first", lambda: _dev(be, grp_first)), G, E + R + G + 1,
                         max_entries=int(max(ent_per_group.max(), 1)),
                         max_comps=int(max(np.diff(grp_first).max(), 1)))
    # ---- stage 2: cluster the class canvases (regions.py:45-68)
    mr2 = int(max(ent_per_group.max(), 1)) + 1
    s2 = ops.cluster_palettes(be, Bm["color"], Bm["off"][:G], Bm["cnt"], None, max_rows=mr2,
                              params=ops.quality_params(be, np.tile(q2, B), mr2))
    fpos2 = ops.first_min(be, Bm["off"], Bm["cnt"], s2["n_leaves"], s2["leaf"], Bm["fpos"])
    # an image with a single non-empty class hands its stage-2 palette to stage 3 unmerged (image.py:246-256)
    grp_regions = np.diff(grp_first).reshape(B, K)
    lone = [(b, int(np.flatnonzero(grp_regions[b])[0])) for b in range(B) if np.count_nonzero(grp_regions[b]) == 1]
    if lone:
        nl2_h, off2_h = s2["n_leaves"].cpu().numpy(), Bm["off"].cpu().numpy()
        for b, k in lone:
            _refuse_duplicate_rows(s2["new_keys"], int(off2_h[b * K + k]), int(nl2_h[b * K + k]), f"stage 2, image {b}")
    # ---- merge level C: classes -> image (image.py:246-256)
    img_first = (np.arange(B + 1) * K).astype(np.int32)
    ent_per_image = (ent_per_group + 1).reshape(B, K).sum(axis=1)
    Cm = ops.merge_level(be, s2["new_keys"], fpos2, Bm["off"], s2["n_leaves"],
                         table.dev(be, "img_first", lambda: _dev(be, img_first)), B,
                         E + R + G + B + 1, max_entries=int(ent_per_image.max()), max_comps=K)
    # ---- stage 3 (image.py:261-286)
    s3 = ops.cluster_palettes(be, Cm["color"], Cm["off"][:B], Cm["cnt"], [q3] * B,
                              max_rows=int(ent_per_image.max()) + 1)
    # ---- compose and paint (last listed class first: merging.py:52)
    group_image_h = (np.arange(G) // K).astype(np.int32)
    ent_final = ops.compose_final(be, P, s1["n_leaves"], ent_off0,
                                  table.dev(be, "seg_region", lambda: _dev(be, table.seg_region)),
                                  table.dev(be, "region_group", lambda: _dev(be, region_group_h)),
                                  table.dev(be, "group_image", lambda: _dev(be, group_image_h)), A, Bm, s2["leaf"],
                                  s2["new_keys"], Cm, s3["leaf"], E)
    out = be.zeros((B, H, W), torch.int16)
    for k in range(K - 1, -1, -1):
        ops.paint(be, labels, crops, ent_off0, ent_final, plane, out, cls=k, idx_bytes=st["idx_bytes"])
    res = EncodeResult(indices=out, palette_keys=s3["new_keys"], palette_off=Cm["off"], palette_cnt=s3["n_leaves"])
    if keep_stages:
        res.stage = dict(st, B=Bm, s2=s2, fpos2=fpos2, C=Cm, s3=s3, ent_final=ent_final, grp_first=grp_first)
    return res


def finish_checks(res: EncodeResult) -> None:
    """Raise if any stage-2/3 problem was refused (reads the small counters back)."""
    ops.check_counts("final palette", res.palette_cnt)


class HostEncoder:
    """The call a host-side user makes: images and label maps in HOST memory in, final palettes and
    index planes in HOST memory out (the boundary of encoder/compression/test.py:100-151, where the
    reference holds numpy arrays and hands `shape` / `palette` / `indices` to the container writer).

    Device and pinned staging buffers are allocated once per table shape (two sets).  `encode` handles
    one batch; `encode_many` streams batches: the host->device copy of batch i+1 runs on a copy stream
    while batch i is being encoded, so a stream of batches runs at max(copy, compute) per batch.
    """

    def __init__(self, be: Backend, table: SegmentTable):
        self.be, self.table = be, table
        B, K, H, W = table.B, table.K, table.H, table.W
        self.cuda = be.device.type == "cuda"
        self.d_img = [be.empty((B, H, W, 3), torch.uint8) for _ in range(2)]
        self.d_lab = [be.empty((K, B, H, W), torch.int32) for _ in range(2)]
        self.h_idx = [torch.empty((B, H, W), dtype=torch.int16, pin_memory=self.cuda) for _ in range(3)]   # results: three slots
        self.h2d_bytes = self.d_img[0].numel() + 4 * self.d_lab[0].numel()
        self.d2h_bytes = 2 * self.h_idx[0].numel()
        self.h_off = [torch.empty((B + 1,), dtype=torch.int32, pin_memory=self.cuda) for _ in range(3)]
        self.h_cnt = [torch.empty((B,), dtype=torch.int32, pin_memory=self.cuda) for _ in range(3)]
        self.h_keys = [torch.empty((B * 4096,), dtype=torch.int32, pin_memory=self.cuda) for _ in range(3)]   # palettes are ~10^2 rows
        if self.cuda:
            self.out_stream = torch.cuda.Stream(device=be.device)
            self.copy_stream = torch.cuda.Stream(device=be.device)
            self.ev_ready = [torch.cuda.Event() for _ in range(2)]
            self.ev_free = [torch.cuda.Event() for _ in range(2)]

    def _upload(self, slot: int, images_host, labels_host, first_use: bool):
        if not self.cuda:
            self.d_img[slot].copy_(images_host); self.d_lab[slot].copy_(labels_host)
            return
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.ev_free[slot])    # the encode that last read this slot is done
            self.d_img[slot].copy_(images_host, non_blocking=True)
            self.d_lab[slot].copy_(labels_host, non_blocking=True)
            self.ev_ready[slot].record(self.copy_stream)

    def _launch(self, slot: int, out: int = 0):
        """Queue the encode of the batch in device slot `slot` and the device->host copies of its results into
        host slot `out` (on the copy-out stream, so that they overlap the next batch's kernels).  Returns the
        handle `_finish` needs."""
        dev = self.be.device
        if self.cuda:
            torch.cuda.current_stream(dev).wait_event(self.ev_ready[slot])
        res = encode_batch(self.be, self.d_img[slot], self.d_lab[slot], self.table)
        B = self.table.B
        if not self.cuda:
            self.h_idx[out].copy_(res.indices)
            return out, res, res.palette_off.clone(), res.palette_cnt.clone(), res.palette_keys, None
        cur = torch.cuda.current_stream(dev)
        self.ev_free[slot].record(cur)
        done = torch.cuda.Event(); done.record(cur)
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(done)
            self.h_idx[out].copy_(res.indices, non_blocking=True)
            self.h_off[out].copy_(res.palette_off[:B + 1], non_blocking=True)
            self.h_cnt[out].copy_(res.palette_cnt[:B], non_blocking=True)
            n_cap = min(res.palette_keys.numel(), self.h_keys[out].numel())
            self.h_keys[out][:n_cap].copy_(res.palette_keys[:n_cap], non_blocking=True)
            out_ev = torch.cuda.Event(); out_ev.record(self.out_stream)
        return out, res, self.h_off[out], self.h_cnt[out], self.h_keys[out], out_ev

    def _finish(self, handle):
        slot, res, off_t, cnt_t, keys_t, out_ev = handle
        if out_ev is not None:
            out_ev.synchronize()
        off, cnt = off_t.cpu().numpy(), cnt_t.cpu().numpy()
        ops.check_counts("final palette", torch.from_numpy(np.ascontiguousarray(cnt)))
        n_keys = int(off[len(cnt) - 1] + cnt[-1]) if len(cnt) else 0
        if n_keys > keys_t.numel():                                 # more palette rows than the staging buffer: fetch directly
            keys_t = res.palette_keys[:n_keys].cpu()
        keys = keys_t[:n_keys].cpu().numpy().astype(np.uint32)
        pals = []
        for b in range(len(cnt)):
            k = keys[off[b]:off[b] + cnt[b]]
            pals.append(np.stack([(k >> 16) & 255, (k >> 8) & 255, k & 255], axis=-1).astype(np.uint8))
        return pals, self.h_idx[slot].numpy().view(np.uint16)

    def _encode_slot(self, slot: int):
        return self._finish(self._launch(slot))

    def encode(self, images_host: torch.Tensor, labels_host: torch.Tensor):
        """images uint8 [B,H,W,3], labels int32 [K,B,H,W] (pinned host tensors for full copy speed).
        Returns (palettes: list of uint8 [m,3] arrays, indices: uint16 [B,H,W] array view of pinned memory)."""
        self._upload(0, images_host, labels_host, first_use=True)
        return self._encode_slot(0)

    def encode_many(self, batches):
        """Generator over (palettes, indices) for an iterable of (images_host, labels_host) batches of the
        table's shape, in order.  Three things overlap: the host->device copy of batch i+1, the encode of batch
        i, and the device->host copy of batch i-1's results.  The index planes are views of three pinned result
        slots used in turn: a result stays valid while the following one is consumed (until next() has been
        called twice more)."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        self._upload(0, cur[0], cur[1], first_use=True)
        i = 0
        pending = None
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                self._upload((i + 1) & 1, nxt[0], nxt[1], first_use=(i == 0))
            handle = self._launch(i & 1, i % 3)                     # blocks the host until this batch's stage 1 is sized
            if pending is not None:
                yield self._finish(pending)                         # batch i-1: its copies ran beside batch i's kernels
            pending = handle
            cur = nxt
            i += 1
        if pending is not None:
            yield self._finish(pending)
