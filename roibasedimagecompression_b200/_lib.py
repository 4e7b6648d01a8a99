"""ctypes binding of librhccq.so (include/rhccq.h).

``lib()`` returns the process-wide backend of the product: the nvcc-built
library next to this file, bound to the current CUDA device.  There is no CPU
path: a missing library, a missing GPU or a GPU that is not sm_100 raises.

``Backend`` itself is parameterised by library path and torch device only so
that the CPU test tier can bind the host-emulation build of the same kernel
sources (tests/emu/build_emu.sh) and run the very same host logic on CPU
tensors.  Nothing in this package constructs such a backend.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librhccq.so")

_P, _I, _Z, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double

# name -> (restype, argtypes); mirrors include/rhccq.h one to one
SIGNATURES = {
    "rhccq_abi_version": (_I, []),
    "rhccq_last_error": (ctypes.c_char_p, []),
    "rhccq_device_check": (_I, []),
    "rhccq_kmeans_rng_fill": (_I, [_P, _I]),
    "rhccq_kmeans_rng_need": (_I, [_I]),
    "rhccq_palette_dbscan_slots": (_I, [_I, _I]),
    "rhccq_workspace_total_bytes": (_Z, [_Z, _I]),
    "rhccq_unique_index_workspace_bytes": (_Z, [_I]),
    "rhccq_palette_dbscan_workspace_bytes": (_Z, [_I, _I]),
    "rhccq_palette_split_workspace_bytes": (_Z, [_I]),
    "rhccq_palette_finish_workspace_bytes": (_Z, [_I]),
    "rhccq_merge_level_workspace_bytes": (_Z, [_I, _I]),
    "rhccq_unique_index": (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "rhccq_cluster_params": (_I, [_P, _P, _I, _P, _P]),
    "rhccq_palette_dbscan": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P, _Z, _P]),
    "rhccq_palette_minibatch_workspace_bytes": (_Z, [_I, _I]),
    "rhccq_palette_minibatch": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _P, _Z, _P]),
    "rhccq_palette_split": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _P, _Z, _P]),
    "rhccq_palette_finish": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "rhccq_remap_first": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _P]),
    "rhccq_merge_level": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "rhccq_first_min": (_I, [_P, _P, _P, _I, _P, _P, _P, _P]),
    "rhccq_compose_final": (_I, [_I] + [_P] * 17),
    "rhccq_paint": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _I, _P, _I, _P, _P]),
    "rhccq_comp_pass": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "rhccq_excl_scan": (_I, [_P, _I, _P, _P]),
    "rhccq_decode_gather": (_I, [_P, _I, ctypes.c_longlong, _P, _I, _P, _P, _P]),
    "rhccq_sq_abs_err": (_I, [_P, _P, ctypes.c_longlong, _P, _P]),
    "rhccq_deflate_chunk_bytes": (_I, []),
    "rhccq_deflate_slot_bytes": (_I, []),
    "rhccq_deflate_chunks": (_I, [_P, ctypes.c_longlong, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rhccq_deflate_pack": (_I, [_P, _P, _P, ctypes.c_longlong, _P, _P]),
    "rhccq_dbscan_plan_make": (_I, [_I, _I, _I, _D, _I, _P, _P, _P]),
    "rhccq_dbscan_workspace_bytes": (_Z, [_P]),
    "rhccq_dbscan_bounds": (_I, [_P, _I, _I, _I, _P, _P, _Z, _P]),
    "rhccq_dbscan_bin": (_I, [_P, _P, _P, _Z, _P]),
    "rhccq_dbscan_count": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_union": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_border": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_relabel": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_flatten": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_attach": (_I, [_P, _P, _Z, _P, _P]),
    "rhccq_dbscan_ws_offset": (_Z, [_P, _I]),
    "rhccq_dbscan_lattice_workspace_bytes": (_Z, [_I, _I]),
    "rhccq_dbscan_lattice_count": (_I, [_P, _I, _I, _I, _D, _I, _P, _P, _P, _P, _Z, _P]),
    "rhccq_dbscan_lattice_union": (_I, [_I, _I, _D, _I, _P, _Z, _P]),
    "rhccq_dbscan_lattice_border": (_I, [_I, _I, _D, _I, _P, _P, _Z, _P]),
    "rhccq_dbscan_lattice_relabel": (_I, [_I, _I, _P, _Z, _P, _P]),
    "rhccq_dbscan_lattice_flatten": (_I, [_I, _I, _P, _P, _Z, _P]),
    "rhccq_dbscan_lattice_attach": (_I, [_I, _I, _D, _I, _P, _Z, _P]),
    "rhccq_dbscan_lattice_ws_offset": (_Z, [_I, _I, _I]),
    "rhccq_uf_own_roots_scratch_ints": (_Z, [_I]),
    "rhccq_uf_own_roots": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "rhccq_uf_merge_edges_gathered": (_I, [_P, _I, _I, _I, _P, _P, _I, _P]),
    "rhccq_uf_rank_labels_gathered": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "rhccq_uf_emit_edges": (_I, [_P, _I, _I, _I, _P, _P, _I, _P]),
    "rhccq_uf_merge_edges": (_I, [_P, _I, _P, _P, _I, _P]),
    "rhccq_uf_lookup_roots": (_I, [_P, _I, _I, _P, _P, _I, _P]),
    "rhccq_dbscan_own_roots": (_I, [_P, _P, _Z, _I, _I, _I, _P, _P, _P]),
    "rhccq_uf_rank_labels": (_I, [_P, _I, _P, _I, _I, _P, _P]),
}


class DbscanPlan(ctypes.Structure):
    """rhccq_dbscan_plan (include/rhccq.h)."""
    _fields_ = [("n", _I), ("dims", _I), ("grid_dims", _I), ("min_pts", _I), ("eps", _D), ("side", _D),
                ("origin", _D * 3), ("ncell", _I * 3), ("n_cells", ctypes.c_longlong), ("cells_per_tile", _I),
                ("n_tiles", ctypes.c_longlong)]


class RhccqError(RuntimeError):
    pass


class Backend:
    """A loaded librhccq build bound to one torch device."""

    def __init__(self, path: str, device):
        if not os.path.exists(path):
            raise RhccqError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.device = torch.device(device)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.cdll, name)          # AttributeError here == ABI mismatch
            fn.restype = res
            fn.argtypes = args
        if self.cdll.rhccq_abi_version() != 1:
            raise RhccqError(f"{path}: ABI version {self.cdll.rhccq_abi_version()} != 1")
        self._rng = None
        self._rng_lock = threading.Lock()
        self.launches = 0                          # kernels launched through this backend (bench.py reports it)
        self._events = None                        # [(name, start, end)] while kernel timing is on (bench.py)

    # -- plumbing ---------------------------------------------------------
    def stream(self) -> int:
        if self.device.type == "cuda":
            return torch.cuda.current_stream(self.device).cuda_stream
        return 0

    def ptr(self, t) -> int:
        if t is None:
            return 0
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a tensor, got {type(t)}")
        if t.device.type != self.device.type:
            raise RhccqError(f"tensor on {t.device}, backend on {self.device}")
        if not t.is_contiguous():
            raise RhccqError("non-contiguous tensor passed to librhccq")
        return t.data_ptr()

    def call(self, name: str, *args, launches: int = 1):
        ev = None
        if self._events is not None and self.device.type == "cuda":
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(self.device))
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise RhccqError(f"{name}: {self.cdll.rhccq_last_error().decode(errors='replace')}")
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            self._events.append((name, ev[0], ev[1]))
        self.launches += launches

    def kernel_timing(self, on: bool) -> None:
        """Bracket every C-ABI launch with CUDA events on the launching stream (bench.py's roofline leg)."""
        self._events = [] if on else None

    def kernel_times_ms(self) -> dict:
        """{entry point: (launches, total ms)} of the launches recorded since kernel_timing(True);
        the caller synchronises first."""
        out: dict = {}
        for name, a, b in self._events or []:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + a.elapsed_time(b))
        return out

    def kernel_trace_ms(self) -> list:
        """[(entry point, ms)] in launch order."""
        return [(name, a.elapsed_time(b)) for name, a, b in self._events or []]

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def workspace(self, need: int, n_problems: int):
        """(tensor or None, bytes) for a per-CTA working set of `need` bytes."""
        total = int(self.cdll.rhccq_workspace_total_bytes(need, n_problems))
        if total == 0:
            return None, 0
        return self.empty((total,), torch.uint8), total

    def rng_table(self, need: int):
        """RandomState(42).random_sample stream on the device, at least `need` long."""
        with self._rng_lock:
            if self._rng is None or self._rng.numel() < need:
                n = max(need, 1 << 16)
                host = torch.empty(n, dtype=torch.float64)
                rc = self.cdll.rhccq_kmeans_rng_fill(host.data_ptr(), n)
                if rc != 0:
                    raise RhccqError(self.cdll.rhccq_last_error().decode())
                self._rng = host.to(self.device)
            return self._rng


_LIB = None
_LIB_LOCK = threading.Lock()


def lib() -> Backend:
    """The CUDA backend (sm_100a).  Raises when it cannot run — never falls back."""
    global _LIB
    with _LIB_LOCK:
        if _LIB is None:
            if not torch.cuda.is_available():
                raise RhccqError("no CUDA device visible: the RHCCQ hot path runs on B200 (sm_100a) only; "
                                 "there is no CPU fallback")
            be = Backend(LIB_PATH, torch.device("cuda", torch.cuda.current_device()))
            if be.cdll.rhccq_device_check() != 0:
                raise RhccqError(be.cdll.rhccq_last_error().decode(errors="replace"))
            _LIB = be
        return _LIB
