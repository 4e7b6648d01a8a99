"""DBSCAN of point clouds on the GPU — the operator one level below the reference's quantiser.

The reference calls ``sklearn.cluster.DBSCAN(eps, min_samples, metric='euclidean').fit_predict(X)``
(/root/reference/encoder/compression/clustering.py:233-235).  ``DBSCAN`` here keeps that operator's
constructor arguments, ``fit`` / ``fit_predict`` and the ``labels_`` / ``core_sample_indices_``
attributes for float32 inputs of 2 to 6 columns, e.g. (x, y, R, G, B) pixel features; labels follow
scikit-learn exactly (clusters numbered by lowest core index, border points in the lowest-numbered
admissible cluster, noise -1).

``dbscan_points`` is the tensor-level form: device tensor in, device tensors out, phases individually
timeable (``phases=`` receives the per-phase CUDA-event times when given).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import Backend, DbscanPlan, RhccqError, lib

PHASES = ("bin", "count", "union", "border", "relabel")


class PointDbscan:
    """A plan + workspace for one (n, dims, eps, min_pts, bounds): reusable across calls on same-sized inputs."""

    def __init__(self, be: Backend, n: int, dims: int, eps: float, min_pts: int, lo, hi, grid_dims: int = 2):
        self.be = be
        self.plan = DbscanPlan()
        lo_a = (ctypes.c_double * 3)(*[float(v) for v in list(lo)[:3]] + [0.0] * (3 - len(list(lo)[:3])))
        hi_a = (ctypes.c_double * 3)(*[float(v) for v in list(hi)[:3]] + [0.0] * (3 - len(list(hi)[:3])))
        rc = be.cdll.rhccq_dbscan_plan_make(int(n), int(dims), int(grid_dims), float(eps), int(min_pts),
                                            ctypes.addressof(lo_a), ctypes.addressof(hi_a), ctypes.addressof(self.plan))
        if rc != 0:
            raise RhccqError("rhccq_dbscan_plan_make: " + be.cdll.rhccq_last_error().decode(errors="replace"))
        self.ws_bytes = int(be.cdll.rhccq_dbscan_workspace_bytes(ctypes.addressof(self.plan)))
        self.ws = be.empty((max(self.ws_bytes, 1),), torch.uint8)
        self.core = be.empty((max(n, 1),), torch.uint8)
        self.labels = be.empty((max(n, 1),), torch.int32)

    def _p(self):
        return ctypes.addressof(self.plan)

    def bin(self, pts):
        self.be.call("rhccq_dbscan_bin", self._p(), self.be.ptr(pts), self.be.ptr(self.ws), self.ws_bytes, self.be.stream(),
                     launches=4)

    def count(self):
        self.be.call("rhccq_dbscan_count", self._p(), self.be.ptr(self.ws), self.ws_bytes, self.be.ptr(self.core),
                     self.be.stream())

    def union(self):
        self.be.call("rhccq_dbscan_union", self._p(), self.be.ptr(self.ws), self.ws_bytes, self.be.ptr(self.core),
                     self.be.stream(), launches=2)

    def border(self):
        self.be.call("rhccq_dbscan_border", self._p(), self.be.ptr(self.ws), self.ws_bytes, self.be.ptr(self.core),
                     self.be.stream(), launches=2)

    def relabel(self):
        self.be.call("rhccq_dbscan_relabel", self._p(), self.be.ptr(self.ws), self.ws_bytes, self.be.ptr(self.labels),
                     self.be.stream(), launches=3)

    def run(self, pts):
        """labels int32 [n], core uint8 [n] (views of buffers owned by the plan)."""
        n = self.plan.n
        if tuple(pts.shape) != (n, self.plan.dims) or pts.dtype != torch.float32:
            raise ValueError(f"points must be float32 [{n},{self.plan.dims}]")
        if n == 0:
            return self.labels[:0], self.core[:0]
        self.bin(pts); self.count(); self.union(); self.border(); self.relabel()
        return self.labels[:n], self.core[:n]


def point_bounds(be: Backend, pts, grid_dims: int = 2):
    """(lo, hi) of the first grid_dims coordinates, computed on the device (one small read-back)."""
    n, dims = pts.shape
    out = be.empty((6,), torch.float64)
    ws = be.empty((1024 * 6 * 4,), torch.uint8)
    be.call("rhccq_dbscan_bounds", be.ptr(pts), int(n), int(dims), int(grid_dims), be.ptr(out), be.ptr(ws), ws.numel(),
            be.stream(), launches=2)
    o = out.cpu().numpy()
    return o[:3], o[3:]


def dbscan_points(be: Backend, pts, eps: float, min_pts: int, grid_dims: int = 2, bounds=None):
    """labels (int32 [n]) and core flags (uint8 [n]) of DBSCAN(eps, min_pts) on float32 points [n, dims]."""
    n, dims = pts.shape
    if n == 0:
        return be.empty((0,), torch.int32), be.empty((0,), torch.uint8)
    lo, hi = bounds if bounds is not None else point_bounds(be, pts, grid_dims)
    plan = PointDbscan(be, n, dims, eps, min_pts, lo, hi, grid_dims)
    labels, core = plan.run(pts)
    return labels.clone(), core.clone()


class DBSCAN:
    """sklearn.cluster.DBSCAN's interface for the euclidean metric (the only one the reference uses)."""

    def __init__(self, eps=0.5, *, min_samples=5, metric="euclidean", grid_dims=None, **unsupported):
        if metric != "euclidean":
            raise NotImplementedError("only metric='euclidean' (the reference's, clustering.py:233)")
        bad = {k: v for k, v in unsupported.items() if k not in ("algorithm", "leaf_size", "n_jobs", "p", "metric_params")}
        if bad:
            raise TypeError(f"unexpected arguments {sorted(bad)}")
        self.eps, self.min_samples, self.metric, self.grid_dims = float(eps), int(min_samples), metric, grid_dims

    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight (the reference never passes it)")
        be = _backend()
        x = np.ascontiguousarray(X, dtype=np.float32)
        if x.ndim != 2 or not 2 <= x.shape[1] <= 6:
            raise ValueError("X must be [n_samples, 2..6]")
        gd = self.grid_dims or (3 if x.shape[1] >= 3 and x.shape[0] > 0 and _dense_in_2d(x, self.eps) else 2)
        labels, core = dbscan_points(be, torch.from_numpy(x).to(be.device), self.eps, self.min_samples, gd)
        self.labels_ = labels.cpu().numpy().astype(np.int64)
        self.core_sample_indices_ = np.flatnonzero(core.cpu().numpy())
        self.components_ = np.asarray(X)[self.core_sample_indices_]
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_


def _dense_in_2d(x: np.ndarray, eps: float) -> bool:
    """True when a 2-D grid of eps cells would hold hundreds of points per cell (then bin on 3 coordinates)."""
    span = (x[:, :2].max(axis=0) - x[:, :2].min(axis=0)) / eps + 1.0
    return x.shape[0] / float(span[0] * span[1]) > 64.0


_BACKEND = None          # tests bind the host-emulation build here


def _backend() -> Backend:
    return _BACKEND if _BACKEND is not None else lib()


# --------------------------------------------------------------------------- one point set across GPUs
def strip_rows(H: int, world: int, rank: int, eps: float, w_xy: float = 1.0):
    """Row ranges of rank `rank` for an H-row raster split into `world` strips.

    Returns (own_r0, own_r1, loc_r0, loc_r1, zone): own rows, the rows the rank reads (own + a halo of
    ceil(2 eps / w_xy) rows per side, clipped), and the list of row ranges of the boundary zone (rows of
    the local block within the halo width of an internal strip boundary) — all in image rows.
    """
    hz = int(np.ceil(2.0 * eps / w_xy))
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    l0, l1 = max(0, r0 - hz), min(H, r1 + hz)
    zone = []
    if rank > 0:
        zone.append((l0, min(r1, r0 + hz)))
    if rank < world - 1:
        lo = max(r0, r1 - hz)
        if zone and lo <= zone[-1][1]:
            zone[-1] = (zone[-1][0], l1)
        else:
            zone.append((lo, l1))
    return r0, r1, l0, l1, zone


class StripDbscan:
    """One rank's share of a DBSCAN over a point set split into strips: the local engine
    (`PointDbscan` for generic points, `LatticeDbscan` for image rows) plus the buffers of the exchange.
    Reusable across calls on same-shaped inputs."""

    def __init__(self, be: Backend, engine, n_loc: int, g0: int, own, zone, group=None):
        self.be, self.engine, self.n_loc, self.g0 = be, engine, int(n_loc), int(g0)
        self.own, self.zone, self.group = (int(own[0]), int(own[1])), [(int(a), int(b)) for a, b in zone], group
        self.lattice = isinstance(engine, LatticeDbscan)
        if self.lattice:
            off = int(be.cdll.rhccq_dbscan_lattice_ws_offset(engine.H, engine.W, 0))
        else:
            off = int(be.cdll.rhccq_dbscan_ws_offset(engine._p(), 0))
        self.rootlab = engine.ws[off:off + 4 * self.n_loc].view(torch.int32)
        self.cap = max(sum(b - a for a, b in self.zone), 1)
        n_own = self.own[1] - self.own[0]
        # The exchange buffers have one fixed capacity on every rank (the largest rank's: one collective here, at
        # construction), their row count travels in the buffer itself, and the kernels skip the padding: a call
        # needs no host synchronisation and allocates nothing.
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        # roots: a cluster has one root; clusters are a small fraction of the points (0.65 % on the synthetic
        # images), room for a sixteenth of the own points + 64 Ki; a rank with more reports it (labels -2, `check`)
        caps = torch.tensor([self.cap, max(n_own // 16 + 65536, 1)], dtype=torch.int64, device=be.device)
        if self.world > 1:
            dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
        self.cap_all, self.ids_all = (int(v) for v in caps.tolist())
        self.e_stride, self.i_stride = 2 + 2 * self.cap_all, 2 + self.ids_all
        self.ebuf = be.zeros((self.e_stride,), torch.int32)          # [0] = edges, rows of (point, root) from int 2
        self.ibuf = be.zeros((self.i_stride,), torch.int32)          # [0] = roots, ascending ids from int 2
        self.eall = be.zeros((self.world, self.e_stride), torch.int32) if self.world > 1 else self.ebuf.view(1, -1)
        self.iall = be.zeros((self.world, self.i_stride), torch.int32) if self.world > 1 else self.ibuf.view(1, -1)
        self.tcap = 1 << max(4, int(np.ceil(np.log2(4 * self.world * self.cap_all))))
        self.tk, self.tp = be.empty((self.tcap,), torch.int32), be.empty((self.tcap,), torch.int32)
        self.scratch = be.empty((int(be.cdll.rhccq_uf_own_roots_scratch_ints(self.n_loc)),), torch.int32)
        self.labels = be.empty((max(n_own, 1),), torch.int32)

    def _all_gather(self, out, mine):
        import torch.distributed as dist
        if self.world == 1:
            return
        if dist.get_backend(self.group) == "gloo":                   # CPU tests
            dist.all_gather([out[r] for r in range(self.world)], mine, group=self.group)
        else:
            dist.all_gather_into_tensor(out, mine, group=self.group)

    def run(self, src, timings: dict | None = None):
        be, eng = self.be, self.engine
        # ---- local phases up to the roots of the local components
        if self.lattice:
            eng.count(src); eng.union()
            be.call("rhccq_dbscan_lattice_flatten", eng.H, eng.W, be.ptr(eng.core), be.ptr(eng.ws), eng.ws_bytes, be.stream())
        else:
            eng.bin(src); eng.count(); eng.union()
            be.call("rhccq_dbscan_flatten", eng._p(), be.ptr(eng.ws), eng.ws_bytes, be.ptr(eng.core), be.stream())
        # ---- boundary edges of this rank (at most one per zone point: they fit by construction), then the one
        # exchange step and the same union-find over all edges on every rank
        self.ebuf[:1].zero_()
        for a, b in self.zone:
            be.call("rhccq_uf_emit_edges", be.ptr(self.rootlab), a, b, self.g0, be.ptr(self.ebuf[2:]), be.ptr(self.ebuf),
                    self.cap_all, be.stream())
        self._all_gather(self.eall, self.ebuf)
        be.call("rhccq_uf_merge_edges_gathered", be.ptr(self.eall), self.world, self.e_stride, self.cap_all, be.ptr(self.tk),
                be.ptr(self.tp), self.tcap, be.stream(), launches=3)
        be.call("rhccq_uf_lookup_roots", be.ptr(self.rootlab), self.n_loc, self.g0, be.ptr(self.tk), be.ptr(self.tp), self.tcap,
                be.stream())
        # ---- border points against global roots, then the global numbering of roots
        if self.lattice:
            be.call("rhccq_dbscan_lattice_attach", eng.H, eng.W, eng.eps, eng.min_pts, be.ptr(eng.ws), eng.ws_bytes, be.stream())
        else:
            be.call("rhccq_dbscan_attach", eng._p(), be.ptr(eng.ws), eng.ws_bytes, be.ptr(eng.core), be.stream())
        be.call("rhccq_uf_own_roots", be.ptr(self.rootlab), self.n_loc, self.own[0], self.own[1], self.g0, be.ptr(self.scratch),
                be.ptr(self.ibuf[2:]), be.ptr(self.ibuf), self.ids_all, be.stream(), launches=3)
        self._all_gather(self.iall, self.ibuf)                       # ascending over the ranks: strips are ordered
        be.call("rhccq_uf_rank_labels_gathered", be.ptr(self.iall), self.world, self.i_stride, self.ids_all, be.ptr(self.rootlab),
                self.own[0], self.own[1], be.ptr(self.labels), be.stream())
        if timings is not None:                                      # (reads counters back: one host synchronisation)
            ne, nr = self.eall[:, 0].cpu(), self.iall[:, 0].cpu()
            if int(nr.max()) > self.ids_all:
                raise RhccqError(f"strip DBSCAN: a rank holds {int(nr.max())} cluster roots, more than the exchange "
                                 f"buffer's {self.ids_all}")
            timings.update(edges_local=int(ne[self.rank]), edges_total=int(ne.sum()), roots_total=int(nr.sum()))
        n_own = self.own[1] - self.own[0]
        return self.labels[:n_own], eng.core[self.own[0]:self.own[1]]

    def check(self) -> None:
        """Raise if the last run overflowed the root exchange buffer (its labels are -2 then)."""
        nr = self.iall[:, 0].cpu()
        if int(nr.max()) > self.ids_all:
            raise RhccqError(f"strip DBSCAN: a rank holds {int(nr.max())} cluster roots, more than the exchange buffer's "
                             f"{self.ids_all}")


def dbscan_strips(be: Backend, pts_local, g0: int, own, zone, eps: float, min_pts: int, group=None, grid_dims: int = 2,
                  timings: dict | None = None):
    """DBSCAN of one point set split into strips, one per rank of `group` (torch.distributed).

    pts_local   float32 [n_loc, dims] on the backend's device: the rank's own points plus the halo, a
                contiguous slice [g0, g0 + n_loc) of the global point order
    own         (lo, hi) local index range of the rank's own points
    zone        list of (lo, hi) local index ranges of the boundary zone (halo and the own points within
                2 eps of an internal boundary)
    Returns labels int32 [own points] identical to the labels an unsplit run gives those points.
    """
    n_loc, dims = pts_local.shape
    lo, hi = point_bounds(be, pts_local, grid_dims)
    eng = PointDbscan(be, n_loc, dims, eps, min_pts, lo, hi, grid_dims)
    return StripDbscan(be, eng, n_loc, g0, own, zone, group).run(pts_local, timings)


def dbscan_image_strips(be: Backend, rows_local, W: int, row0: int, own_rows, zone_rows, eps: float, min_pts: int,
                        group=None, timings: dict | None = None):
    """The same for image rows through the lattice kernels.  rows_local: uint8 [h_loc, W, 3] (or float32
    points [h_loc * W, 5]) = image rows [row0, row0 + h_loc); own_rows / zone_rows in local row indices."""
    h_loc = rows_local.shape[0] if rows_local.dtype == torch.uint8 else rows_local.shape[0] // W
    eng = LatticeDbscan(be, h_loc, W, eps, min_pts)
    st = StripDbscan(be, eng, h_loc * W, row0 * W, (own_rows[0] * W, own_rows[1] * W),
                     [(a * W, b * W) for a, b in zone_rows], group)
    return st.run(rows_local, timings)


# --------------------------------------------------------------------------- image lattice fast path
class LatticeDbscan:
    """DBSCAN of the pixel features (x, y, R, G, B) of an H x W image (w_xy = 1) through the lattice
    kernels: stencil + packed 8-bit colours instead of cell binning.  Same labels as `PointDbscan`."""

    def __init__(self, be: Backend, H: int, W: int, eps: float, min_pts: int):
        self.be, self.H, self.W, self.eps, self.min_pts = be, int(H), int(W), float(eps), int(min_pts)
        n = self.H * self.W
        self.ws_bytes = int(be.cdll.rhccq_dbscan_lattice_workspace_bytes(self.H, self.W))
        self.ws = be.empty((self.ws_bytes,), torch.uint8)
        self.count_out = be.empty((n,), torch.int32)
        self.core = be.empty((n,), torch.uint8)
        self.labels = be.empty((n,), torch.int32)
        self.status = be.zeros((1,), torch.int32)

    def count(self, src):
        """src: float32 points [H*W, 5] or uint8 image [H, W, 3] on the device."""
        kind = 0 if src.dtype == torch.float32 else 1
        if kind == 0 and tuple(src.shape) != (self.H * self.W, 5):
            raise ValueError("points must be float32 [H*W, 5]")
        if kind == 1 and (src.dtype != torch.uint8 or tuple(src.shape) != (self.H, self.W, 3)):
            raise ValueError("image must be uint8 [H, W, 3]")
        self.be.call("rhccq_dbscan_lattice_count", self.be.ptr(src), kind, self.H, self.W, self.eps, self.min_pts,
                     self.be.ptr(self.count_out), self.be.ptr(self.core), self.be.ptr(self.status), self.be.ptr(self.ws),
                     self.ws_bytes, self.be.stream())

    def union(self):
        self.be.call("rhccq_dbscan_lattice_union", self.H, self.W, self.eps, self.min_pts, self.be.ptr(self.ws), self.ws_bytes,
                     self.be.stream(), launches=2)

    def border(self):
        self.be.call("rhccq_dbscan_lattice_border", self.H, self.W, self.eps, self.min_pts, self.be.ptr(self.core),
                     self.be.ptr(self.ws), self.ws_bytes, self.be.stream(), launches=2)

    def relabel(self):
        self.be.call("rhccq_dbscan_lattice_relabel", self.H, self.W, self.be.ptr(self.ws), self.ws_bytes,
                     self.be.ptr(self.labels), self.be.stream(), launches=3)

    def run(self, src, check: bool = True):
        """labels int32 [H*W], core uint8 [H*W].  With float32 points the kernel verifies that they are the
        lattice of an 8-bit image; `check` reads that flag back and raises if they are not."""
        self.count(src); self.union(); self.border(); self.relabel()
        if check and src.dtype == torch.float32 and int(self.status.item()) != 0:
            raise RhccqError("points are not the pixel lattice of an 8-bit image (x = column, y = row, integer "
                             "colours 0..255): use dbscan_points")
        return self.labels, self.core


def dbscan_image(be: Backend, image_u8, eps: float, min_pts: int):
    """Labels [H, W] and core flags of DBSCAN over the (x, y, R, G, B) features of an image on the device."""
    H, W, _ = image_u8.shape
    plan = LatticeDbscan(be, H, W, eps, min_pts)
    labels, core = plan.run(image_u8)
    return labels.view(H, W).clone(), core.view(H, W).clone()
