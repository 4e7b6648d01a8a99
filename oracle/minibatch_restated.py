"""Exact-arithmetic restatement of the MiniBatchKMeans call of the reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Palettes of 10 000 or more non-black colours are not clustered by DBSCAN but by
``MiniBatchKMeans(n_clusters=ceil(n*q/100/10), batch_size=1000, random_state=42,
n_init='auto').fit_predict(colours.astype(float))``
(/root/reference/encoder/compression/clustering.py:207-218).  That is scikit-learn
(requirements.txt:6, unpinned; 1.9.0 here): sklearn/cluster/_kmeans.py:2056-2229 (fit),
:1566-1684 (_mini_batch_step), :1974-2037 (_mini_batch_convergence), :2039-2054 (_random_reassign),
:964-1043 (_init_centroids), :180-282 (_kmeans_plusplus), _k_means_minibatch.pyx:68-118 (centre update).

The float evaluation follows scikit-learn operation by operation, as oracle/kmeans_sklearn.c does for KMeans:

* labels (every batch, and the final prediction over the whole palette) are those of ``_labels_inertia``: the
  chunked E step, ``|c|^2 - 2 x.c`` through dgemm in chunks of 256 samples, first minimum
  (``kmeans_restated.estep_labels`` -> oracle/kmeans_sklearn.c: e_step; MiniBatchKMeans does not centre the data);
* the distance of a point to its centre that enters the batch inertia is ``_inertia_dense``'s direct
  ``((x0-c0)^2 + (x1-c1)^2) + (x2-c2)^2``; the inertia is their sum in batch order (sklearn: an OpenMP reduction
  whose order depends on the thread count; it only feeds the early-stopping average);
* the k-means++ sums are exact integers (uncentred uint8 data: every term of sklearn's float evaluation is an
  integer below 2^53, so its results are the same integers);
* the centre update is the scalar loop of _k_means_minibatch.pyx, in batch order;
* ``RandomState(42)`` is consumed by numpy's own ``randint`` / ``choice`` / ``random_sample`` in sklearn's order.

With these, labels AND final centres equal scikit-learn's bit for bit (tests/test_oracle_golden.py) on every
palette on which one branch is not taken: when more than half a batch of centres would be reassigned
(``to_reassign.sum() > 0.5 * batch_size``, which needs k > 500 and in practice k > ~1000), scikit-learn keeps the
centres with the largest counts by ``np.argsort(weight_sums)``, numpy's default UNSTABLE sort, and the counts are
small integers full of ties.  Which tied centres are kept then depends on numpy's CPU dispatch (x86-simd-sort's
AVX-512 argsort, its AVX2 one and the scalar introsort order the ties differently: the same call gives three
different answers under NPY_DISABLE_CPU_FEATURES), so the reference's own output is machine-dependent there.
This restatement — and the CUDA kernel — keep the STABLE order (ties by index); ``info["unstable_cuts"]`` counts
how often a call went through that branch.

The CUDA kernel (csrc/rhccq_minibatch.cu) must match this file bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

BATCH_SIZE = 1000            # clustering.py:213
MAX_ITER = 100               # MiniBatchKMeans defaults
MAX_NO_IMPROVEMENT = 10
REASSIGNMENT_RATIO = 0.01


def n_clusters_for(n_colors: int, quality: float) -> int:
    """clustering.py:210."""
    return math.ceil(n_colors * (quality / 100) / 10)


def _d2_int(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    d = x - c
    return (d * d).sum(axis=-1)


def _kmeans_pp(xs: np.ndarray, k: int, rs: np.random.RandomState) -> np.ndarray:
    """_kmeans_plusplus (sklearn/_kmeans.py:216-282) on integer colours; returns seed rows of xs."""
    n = xs.shape[0]
    t = 2 + int(math.log(k))
    w = np.ones(n)
    first = int(rs.choice(n, p=w / w.sum()))                 # :232
    seeds = np.empty(k, dtype=np.int64)
    seeds[0] = first
    closest = _d2_int(xs, xs[first])
    pot = int(closest.sum())
    for c in range(1, k):
        rv = rs.uniform(size=t) * float(pot)                 # :247
        cum = np.cumsum(closest)
        cand = np.searchsorted(cum.astype(np.float64), rv, side="left")
        np.clip(cand, None, n - 1, out=cand)
        dist = _d2_int(xs[None, :, :], xs[cand][:, None, :])
        np.minimum(dist, closest[None, :], out=dist)
        pots = dist.sum(axis=1)
        b = int(np.argmin(pots))
        pot = int(pots[b])
        closest = dist[b]
        seeds[c] = cand[b]
    return seeds


def _nearest(xf: np.ndarray, centers: np.ndarray):
    """(labels, own distance) as sklearn's _labels_inertia returns them: the labels by the chunked E step
    (``|c|^2 - 2 x.c`` through dgemm, first minimum: oracle/kmeans_sklearn.c), the distance of a point to its
    centre by _inertia_dense's direct ``((d0^2 + d1^2) + d2^2)``."""
    from . import kmeans_restated as K
    lab = K.estep_labels(xf, centers)
    d = xf - centers[lab]
    own = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return lab, own


def minibatch_labels(colors: np.ndarray, k: int, return_info: bool = False):
    """Labels of ``MiniBatchKMeans(k, batch_size=1000, random_state=42, n_init='auto').fit_predict``."""
    x = np.asarray(colors).astype(np.int64).reshape(-1, 3)
    n = x.shape[0]
    if not 1 <= k <= n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    xf = x.astype(np.float64)
    rs = np.random.RandomState(42)
    batch = min(BATCH_SIZE, n)                               # :1935
    init_size = 3 * batch                                    # :1938-1942
    if init_size < k:
        init_size = 3 * k
    init_size = min(init_size, n)                            # :1954
    rs.randint(0, n, init_size)                              # :2110 validation set (only consumed: n_init == 1)
    if init_size < n:                                        # :1012-1017
        xs = x[rs.randint(0, n, init_size)]
    else:
        xs = x
    centers = xs[_kmeans_pp(xs, k, rs)].astype(np.float64)
    centers_new = np.empty_like(centers)
    counts = np.zeros(k, dtype=np.float64)
    ewa = ewa_min = None
    no_improvement = 0
    n_since = 0
    n_steps = (MAX_ITER * n) // batch                        # :2163
    p = np.ones(n) / float(n)                                # :2164 normalized_sample_weight (sum of ones == n)
    steps_done = 0
    unstable_cuts = 0
    for step in range(n_steps):
        idx = rs.choice(n, batch, p=p, replace=True)         # :2171-2176
        xb, xbf = x[idx], xf[idx]
        n_since += batch                                     # :2047-2054
        reassign = bool((counts == 0).any() or n_since >= 10 * k)
        if reassign:
            n_since = 0
        lab, own = _nearest(xbf, centers)
        inertia = 0.0
        for v in own.tolist():                               # batch order
            inertia = inertia + v
        for j in range(k):                                   # _k_means_minibatch.pyx:68-118
            members = np.flatnonzero(lab == j)
            if members.size:
                c = centers[j] * counts[j]
                for i in members:
                    c = c + xbf[i]
                counts[j] = counts[j] + float(members.size)
                centers_new[j] = c * (1.0 / counts[j])
            else:
                centers_new[j] = centers[j]
        if reassign:                                         # :1652-1682
            to_reassign = counts < REASSIGNMENT_RATIO * counts.max()
            if to_reassign.sum() > 0.5 * batch:
                unstable_cuts += 1
                keep = np.argsort(counts, kind="stable")[int(0.5 * batch):]
                to_reassign[keep] = False
            n_re = int(to_reassign.sum())
            if n_re:
                new = rs.choice(batch, replace=False, size=n_re)
                centers_new[to_reassign] = xbf[new]
            if (~to_reassign).any():
                counts[to_reassign] = np.min(counts[~to_reassign])
        centers, centers_new = centers_new, centers          # :2203
        steps_done = step + 1
        # _mini_batch_convergence (:1974-2037); tol == 0 for MiniBatchKMeans defaults
        bi = inertia / batch
        if step == 0:
            continue
        if ewa is None:
            ewa = bi
        else:
            alpha = min(batch * 2.0 / (n + 1), 1.0)
            ewa = ewa * (1 - alpha) + bi * alpha
        if ewa_min is None or ewa < ewa_min:
            no_improvement = 0
            ewa_min = ewa
        else:
            no_improvement += 1
        if no_improvement >= MAX_NO_IMPROVEMENT:
            break
    labels, _ = _nearest(xf, centers)
    if return_info:
        return labels, {"steps": steps_done, "centers": centers, "counts": counts, "unstable_cuts": unstable_cuts}
    return labels
