"""Exact-arithmetic restatement of the K-Means the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference splits over-large colour clusters with
``KMeans(n_clusters=k, random_state=42, n_init='auto').fit_predict(colours.astype(float))``
(/root/reference/encoder/compression/clustering.py:751-752).  That is
scikit-learn (requirements.txt:6, unpinned; 1.9.0 in this image): one
k-means++ seeding drawn from ``RandomState(42)`` followed by Lloyd iterations
(sklearn/cluster/_kmeans.py:180-282 seeding, :626-757 Lloyd,
_k_means_lloyd.pyx:_update_chunk_dense E/M step,
_k_means_common.pyx:167-211 empty-cluster relocation).

scikit-learn evaluates this in float64 with BLAS and an OpenMP reduction whose
rounding depends on the CPU kernel and thread count, so its labels are not
reproducible bit for bit even between two hosts.  This restatement keeps the
algorithm and the random stream and replaces the rounding-dependent pieces by
exact ones, so that a second implementation (the CUDA kernel) can match it
exactly:

* inputs are uint8 colours, so seeding distances are integers and the
  potential / cumulative sums are exact int64 (sklearn: float64 of the same
  numbers, rounded);
* a centre is ``S_j / n_j`` with ``S_j`` the exact integer sum of its members
  (sklearn: float sum of mean-centred members times a reciprocal);
* a point-to-centre distance is ``((x0-c0)^2 + (x1-c1)^2) + (x2-c2)^2`` in
  IEEE double with no fused multiply-add, first minimum wins
  (sklearn: ``|c|^2 - 2 x.c`` through GEMM on centred data);
* the tolerance and the centre shift are evaluated in a fixed order.

Agreement with scikit-learn itself is measured in tests/test_oracle_kmeans.py
(identical seeding; labels identical or differing only at exact distance ties).
"""
from __future__ import annotations

import math

import numpy as np

MAX_ITER = 300          # sklearn KMeans default max_iter
TOL = 1e-4              # sklearn KMeans default tol

_RNG_CACHE: dict[int, np.ndarray] = {}


def rng_doubles(count: int) -> np.ndarray:
    """First ``count`` outputs of ``RandomState(42).random_sample``.

    Every K-Means call of the reference builds a fresh ``RandomState(42)``
    (clustering.py:751), so all calls consume the same stream.
    """
    have = _RNG_CACHE.get(42)
    if have is None or have.size < count:
        have = np.random.RandomState(42).random_sample(max(count, 4096))
        _RNG_CACHE[42] = have
    return have[:count]


def n_local_trials(k: int) -> int:
    """sklearn/cluster/_kmeans.py:226  ``2 + int(np.log(n_clusters))``."""
    return 2 + int(math.log(k))


def _d2_int(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    d = x - c
    return (d * d).sum(axis=-1)


def kmeans_pp_seeds(colors: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k-means++ seeds (sklearn/_kmeans.py:216-282), int-exact."""
    x = np.asarray(colors).astype(np.int64).reshape(-1, 3)
    n = x.shape[0]
    t = n_local_trials(k)
    r = rng_doubles(1 + (k - 1) * t)
    # :229  random_state.choice(n, p=uniform)  ==  floor(u * n) for u = r[0]
    first = min(int(r[0] * n), n - 1)
    seeds = np.empty(k, dtype=np.int64)
    seeds[0] = first
    closest = _d2_int(x, x[first])                       # :238-240
    pot = int(closest.sum())                             # :241
    ri = 1
    for c in range(1, k):
        rv = r[ri:ri + t] * float(pot)                   # :247
        ri += t
        cum = np.cumsum(closest)                         # :248-250 (exact int64)
        cand = np.searchsorted(cum.astype(np.float64), rv, side="left")
        np.clip(cand, None, n - 1, out=cand)             # :252
        dist = _d2_int(x[None, :, :], x[cand][:, None, :])   # [t, n]
        np.minimum(dist, closest[None, :], out=dist)     # :260
        pots = dist.sum(axis=1)                          # :261
        b = int(np.argmin(pots))                         # :264 first minimum
        pot = int(pots[b])
        closest = dist[b]
        seeds[c] = cand[b]
    return seeds


def tolerance(x_int: np.ndarray) -> float:
    """``mean(var(X, axis=0)) * tol`` (sklearn/_kmeans.py:285-293), fixed order."""
    n = x_int.shape[0]
    s1 = x_int.sum(axis=0)
    s2 = (x_int * x_int).sum(axis=0)
    num = n * s2 - s1 * s1                               # exact int64, >= 0
    nn = float(n) * float(n)
    v = [float(int(num[d])) / nn for d in range(3)]
    return ((v[0] + v[1]) + v[2]) / 3.0 * TOL


def _e_step(xf: np.ndarray, centers: np.ndarray):
    """Labels (first minimum) and the distance of each point to its own centre."""
    n, k = xf.shape[0], centers.shape[0]
    labels = np.empty(n, dtype=np.int64)
    own = np.empty(n, dtype=np.float64)
    step = max(1, (1 << 22) // max(k, 1))                # bound the [rows, k] temporaries
    for lo in range(0, n, step):
        xs = xf[lo:lo + step]
        d0 = xs[:, None, 0] - centers[None, :, 0]
        d1 = xs[:, None, 1] - centers[None, :, 1]
        d2 = xs[:, None, 2] - centers[None, :, 2]
        dist = (d0 * d0 + d1 * d1) + d2 * d2
        lab = np.argmin(dist, axis=1)                    # first minimum
        labels[lo:lo + step] = lab
        own[lo:lo + step] = dist[np.arange(xs.shape[0]), lab]
    return labels, own


def kmeans_labels(colors: np.ndarray, k: int, return_info: bool = False):
    """Labels of ``KMeans(k, random_state=42, n_init='auto').fit_predict``.

    ``colors`` is ``uint8 [n,3]`` in the order the reference passes them
    (ascending palette index).  ``k <= n`` is required (sklearn raises
    otherwise; the reference clamps at clustering.py:742).
    """
    x = np.asarray(colors).astype(np.int64).reshape(-1, 3)
    n = x.shape[0]
    if not 1 <= k <= n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    xf = x.astype(np.float64)
    seeds = kmeans_pp_seeds(x, k)
    centers = xf[seeds].copy()
    tol = tolerance(x)
    s_all = x.sum(axis=0)
    labels_old = np.full(n, -1, dtype=np.int64)
    labels = labels_old
    strict = False
    n_iter = 0
    for it in range(MAX_ITER):
        n_iter = it + 1
        labels, own = _e_step(xf, centers)
        cnt = np.bincount(labels, minlength=k).astype(np.int64)
        sums = np.zeros((k, 3), dtype=np.int64)
        np.add.at(sums, labels, x)
        empties = np.flatnonzero(cnt == 0)
        if empties.size:                                 # _k_means_common.pyx:177-211
            if own.max() != 0:
                # farthest points first; ties by lower index (argpartition order
                # in sklearn is unspecified among ties)
                far = np.lexsort((np.arange(n), -own))[:empties.size]
                for e, f in zip(empties, far):
                    old = labels[f]
                    sums[old] -= x[f]
                    sums[e] = x[f]
                    cnt[e] = 1
                    cnt[old] -= 1
        new_centers = np.empty_like(centers)
        for j in range(k):
            if cnt[j] > 0:
                c = float(cnt[j])
                new_centers[j] = [float(int(sums[j, d])) / c for d in range(3)]
            else:                                        # un-relocated empty cluster:
                new_centers[j] = [float(int(s_all[d])) / float(n) for d in range(3)]
        shift = 0.0
        for j in range(k):                               # fixed order
            a = new_centers[j] - centers[j]
            shift = shift + ((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])
        centers = new_centers
        if np.array_equal(labels, labels_old):           # _kmeans.py:717-722
            strict = True
            break
        if shift <= tol:                                 # :724-733
            break
        labels_old = labels
    if not strict:                                       # :737-749
        labels, _ = _e_step(xf, centers)
    if return_info:
        return labels.astype(np.int64), {"n_iter": n_iter, "strict": strict,
                                         "seeds": seeds, "centers": centers}
    return labels.astype(np.int64)
