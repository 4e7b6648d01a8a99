"""Restatement of the K-Means the reference calls, in scikit-learn's own arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference splits over-large colour clusters with
``KMeans(n_clusters=k, random_state=42, n_init='auto').fit_predict(colours.astype(float))``
(/root/reference/encoder/compression/clustering.py:751-752): scikit-learn (requirements.txt:6,
unpinned; 1.9.0 in this image), i.e. one k-means++ seeding drawn from ``RandomState(42)`` followed by
Lloyd iterations in float64 on mean-centred data.  Which colour goes to which cluster depends on the
rounding of every step, so the restatement follows scikit-learn operation by operation, including the
summation orders of the numpy / OpenBLAS kernels underneath: ``oracle/kmeans_sklearn.c`` (the header of
that file lists every formula and the scikit-learn lines it follows).  This module loads the compiled
file (``make -C oracle``; ``__graft_entry__.build()`` runs it) and feeds it the random stream.

Pinned: ``tools/kmeans_replay.py`` replays every K-Means call recorded from the reference itself by
``tests/golden/make_golden.py`` (3 693 calls) and compares label for label — all identical
(``tests/golden/kmeans_replay.json``); ``tests/test_oracle_golden.py`` re-checks the recorded calls on
every run and compares with scikit-learn itself on random palettes where it is installed.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkmeans_sklearn.so")
_LIB = None
_RNG_CACHE: dict[int, np.ndarray] = {}


def _lib():
    global _LIB
    if _LIB is None:
        src = os.path.join(_HERE, "kmeans_sklearn.c")
        if not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = ctypes.CDLL(_SO)
        lib.km_sklearn_labels.restype = ctypes.c_int
        lib.km_sklearn_labels.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.km_first_seed.restype = ctypes.c_int
        lib.km_first_seed.argtypes = [ctypes.c_int, ctypes.c_double]
        lib.km_gemm_edge_rows.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.km_estep_labels.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.km_estep_labels.restype = None
        _LIB = lib
    return _LIB


def rng_doubles(count: int) -> np.ndarray:
    """First ``count`` outputs of ``RandomState(42).random_sample``.

    Every K-Means call of the reference builds a fresh ``RandomState(42)`` (clustering.py:751), so all
    calls consume the same stream: one draw for the first centre (``choice``), then
    ``uniform(size=n_local_trials)`` per further centre (sklearn/cluster/_kmeans.py:229,247).
    """
    have = _RNG_CACHE.get(42)
    if have is None or have.size < count:
        have = np.random.RandomState(42).random_sample(max(count, 1 << 16))
        _RNG_CACHE[42] = have
    return have[:count]


def n_local_trials(k: int) -> int:
    """sklearn/cluster/_kmeans.py:226  ``2 + int(np.log(n_clusters))``."""
    return 2 + int(math.log(k))


def first_seed(n: int) -> int:
    """Index ``RandomState(42).choice(n, p=uniform)`` returns (sklearn/cluster/_kmeans.py:229)."""
    return int(_lib().km_first_seed(int(n), float(rng_doubles(1)[0])))


def gemm_edge_rows(k: int) -> tuple[int, int]:
    """Cluster rows [lo, hi) of Lloyd's dgemm that the 4-row edge kernel computes (kmeans_sklearn.c)."""
    lo, hi = ctypes.c_int(0), ctypes.c_int(0)
    _lib().km_gemm_edge_rows(int(k), ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


def kmeans_labels(colors: np.ndarray, k: int, return_info: bool = False):
    """Labels of ``KMeans(k, random_state=42, n_init='auto').fit_predict(colors.astype(float))``.

    ``colors`` is ``uint8 [n,3]`` in the order the reference passes them (ascending palette index).
    ``k <= n`` is required (scikit-learn raises otherwise; the reference clamps at clustering.py:742).
    """
    c = np.ascontiguousarray(np.asarray(colors).astype(np.uint8).reshape(-1, 3))
    n = c.shape[0]
    k = int(k)
    if not 1 <= k <= n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    r = np.ascontiguousarray(rng_doubles(1 + (k - 1) * n_local_trials(k)))
    labels = np.empty(n, dtype=np.int32)
    seeds = np.empty(k, dtype=np.int64)
    info = np.zeros(3, dtype=np.int32)
    rc = _lib().km_sklearn_labels(c.ctypes.data, n, k, r.ctypes.data, labels.ctypes.data, seeds.ctypes.data,
                                  info.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"km_sklearn_labels failed ({rc})")
    if return_info:
        return labels.astype(np.int64), {"n_iter": int(info[0]), "strict": bool(info[1]),
                                         "relocations": int(info[2]), "seeds": seeds}
    return labels.astype(np.int64)


def kmeans_pp_seeds(colors: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k-means++ seeds (sklearn/cluster/_kmeans.py:216-282)."""
    return kmeans_labels(colors, k, return_info=True)[1]["seeds"]


def estep_labels(x: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """Labels of float64 points x[n,3] against centers[k,3] by scikit-learn's chunked E step (_labels_inertia):
    ``|c|^2 - 2 x.c`` through dgemm in chunks of 256 samples, first minimum."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
    lab = np.empty(len(x), dtype=np.int32)
    _lib().km_estep_labels(x.ctypes.data, len(x), c.ctypes.data, len(c), lab.ctypes.data)
    return lab.astype(np.int64)
