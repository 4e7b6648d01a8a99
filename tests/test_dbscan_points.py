"""Large-N point DBSCAN (bin / count / union / border / relabel) against scikit-learn's golden labels
and the brute-force oracle.  Labels must be IDENTICAL to scikit-learn's (cluster numbering included)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import rhccq_oracle as O
from roibasedimagecompression_b200 import dbscan as D
from roibasedimagecompression_b200.synth import synth, pixel_features


def _run(be, pts, eps, mp, gd=2):
    lab, core = D.dbscan_points(be, torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(be.device), eps, mp, gd)
    return lab.cpu().numpy().astype(np.int64), core.cpu().numpy().astype(bool)


def test_points_golden_sklearn(backend):
    g = golden("dbscan_points.npz")
    pts = g["points"]
    for i, (eps, mp) in enumerate(g["combos"]):
        lab, core = _run(backend, pts, float(eps), int(mp))
        assert np.array_equal(lab, g[f"labels{i}"]), (eps, mp)
        assert np.array_equal(core, g[f"core{i}"]), (eps, mp)
    u = g["uniform_points"]
    for i, (eps, mp) in enumerate(g["uniform_combos"]):
        for gd in (2, 3):
            lab, _ = _run(backend, u, float(eps), int(mp), gd)
            assert np.array_equal(lab, g[f"uniform_labels{i}"]), (eps, mp, gd)


def test_points_ties_ragged_and_empty(backend):
    # integer lattice: many pairs at exactly eps (float64 tie path), eps^2 integer and not
    yy, xx = np.mgrid[0:23, 0:37]
    pts = np.stack([xx.ravel(), yy.ravel(), (xx.ravel() * 7 + yy.ravel() * 3) % 5], axis=1).astype(np.float32)
    for eps, mp in ((1.0, 3), (2.0, 5), (np.sqrt(2.0), 4), (2.2360679774997896, 9), (3.0, 40)):
        lab, core = _run(backend, pts, float(eps), mp)
        want = O.dbscan_labels(pts, float(eps), mp)
        assert np.array_equal(lab, want), (eps, mp)
    lab, core = _run(backend, np.zeros((0, 5), np.float32), 1.0, 2)
    assert lab.size == 0 and core.size == 0
    one = np.array([[3.0, 4.0, 5.0]], np.float32)
    assert _run(backend, one, 1.0, 1)[0].tolist() == [0] and _run(backend, one, 1.0, 2)[0].tolist() == [-1]
    same = np.tile(np.array([[1.0, 2.0, 3.0, 4.0, 5.0]], np.float32), (700, 1))       # one cell, many chunks
    lab, core = _run(backend, same, 0.5, 700)
    assert (lab == 0).all() and core.all()


def test_points_sklearn_interface(backend):
    D._BACKEND = backend if backend.device.type == "cpu" else None
    try:
        img = synth(24, 32, 9)
        pts = pixel_features(img)
        m = D.DBSCAN(eps=4.0, min_samples=4).fit(pts)
        want = O.dbscan_labels(pts, 4.0, 4)
        assert np.array_equal(m.labels_, want)
        assert np.array_equal(D.DBSCAN(eps=4.0, min_samples=4).fit_predict(pts), want)
        assert m.labels_.dtype == np.int64 and m.components_.shape[1] == 5
    finally:
        D._BACKEND = None


@pytest.mark.gpu
def test_points_large_properties():
    """4K-image-sized input: properties that need no O(n^2) oracle — idempotent canonical labels, every
    core point's neighbours inside its cluster on a sampled subset, cluster ids dense and ordered."""
    from roibasedimagecompression_b200._lib import lib
    be = lib()
    img = synth(1080, 1920, 1234)
    pts = pixel_features(img)
    lab, core = _run(be, pts, 3.0, 8)
    lab2, core2 = _run(be, pts, 3.0, 8)
    assert np.array_equal(lab, lab2) and np.array_equal(core, core2)             # deterministic despite atomics
    k = lab.max() + 1
    first = np.full(k, len(lab), np.int64)
    cl = np.flatnonzero(core)
    np.minimum.at(first, lab[cl], cl)
    assert (np.diff(first) > 0).all()                                            # numbered by lowest core index
    # window check: brute force inside 40x40 pixel windows (interior points only)
    H, W = 1080, 1920
    rng = np.random.default_rng(0)
    for _ in range(6):
        r0, c0 = int(rng.integers(0, H - 40)), int(rng.integers(0, W - 40))
        rows, cols = np.mgrid[r0:r0 + 40, c0:c0 + 40]
        ids = (rows * W + cols).ravel()
        sub = pts[ids].astype(np.float64)
        d2 = ((sub[:, None, :] - sub[None, :, :]) ** 2).sum(-1)
        inner = ((rows > r0 + 3) & (rows < r0 + 36) & (cols > c0 + 3) & (cols < c0 + 36)).ravel()
        cnt = (d2 <= 9.0).sum(1)
        assert np.array_equal(core[ids][inner], (cnt >= 8)[inner])
        ci = np.flatnonzero(inner & core[ids])
        for a in ci[:200]:
            nb = np.flatnonzero((d2[a] <= 9.0) & core[ids])
            assert (lab[ids][nb] == lab[ids][a]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("eps,mp", [(3.0, 8), (2.0, 4), (5.0, 16)])
def test_points_quarter_million_vs_sklearn(eps, mp):
    """Deep union-find trees and multi-block scans only show up at scale: 262 144 points against
    scikit-learn itself (the operator the reference calls), labels identical."""
    sk = pytest.importorskip("sklearn.cluster")
    from roibasedimagecompression_b200._lib import lib
    be = lib()
    pts = pixel_features(synth(512, 512, 77))
    lab, core = _run(be, pts, eps, mp)
    ref = sk.DBSCAN(eps=eps, min_samples=mp).fit(pts)
    rc = np.zeros(len(pts), bool)
    rc[ref.core_sample_indices_] = True
    assert np.array_equal(core, rc)
    assert np.array_equal(lab, ref.labels_)


# --------------------------------------------------------------------------- image lattice fast path
def test_lattice_golden_sklearn(backend):
    g = golden("dbscan_points.npz")
    img, pts = g["image"], g["points"]
    H, W, _ = img.shape
    be = backend
    for i, (eps, mp) in enumerate(g["combos"]):
        lab, core = D.dbscan_image(be, torch.from_numpy(img.copy()).to(be.device), float(eps), int(mp))
        assert np.array_equal(lab.cpu().numpy().reshape(-1), g[f"labels{i}"]), (eps, mp)
        assert np.array_equal(core.cpu().numpy().reshape(-1).astype(bool), g[f"core{i}"]), (eps, mp)
        plan = D.LatticeDbscan(be, H, W, float(eps), int(mp))          # the float32-point entry of the same kernels
        lab2, _ = plan.run(torch.from_numpy(pts).to(be.device))
        assert np.array_equal(lab2.cpu().numpy(), g[f"labels{i}"])


def test_lattice_rejects_non_lattice_points_and_ragged_sizes(backend):
    be = backend
    img = synth(19, 67, 5)                                              # not a multiple of the 64 x 16 tile
    pts = pixel_features(img)
    for eps, mp in ((1.0, 2), (2.5, 5), (4.0, 30)):
        lab, core = D.dbscan_image(be, torch.from_numpy(img.copy()).to(be.device), eps, mp)
        assert np.array_equal(lab.cpu().numpy().reshape(-1), O.dbscan_labels(pts, eps, mp)), (eps, mp)
    tiny = synth(5, 9, 3)                                               # fewer pixels than the relabel's bit tables need
    tiny[:, :4] = tiny[0, 0]
    lab, _ = D.dbscan_image(be, torch.from_numpy(tiny.copy()).to(be.device), 2.0, 3)
    assert np.array_equal(lab.cpu().numpy().reshape(-1), O.dbscan_labels(pixel_features(tiny), 2.0, 3))
    bad = pts.copy()
    bad[100, 2] += 0.5
    plan = D.LatticeDbscan(be, 19, 67, 2.0, 3)
    with pytest.raises(Exception):
        plan.run(torch.from_numpy(bad).to(be.device))


def test_lattice_float_points_every_small_radius(backend):
    """float32 [n,5] points of an image whose width is a multiple of 4 but not of the 64-pixel tile: on the GPU this
    is the TMA-fed count kernel (compile-time stencil for integer eps, run-time otherwise) and the row-window union;
    several tiles in both directions, partial tiles on the right and bottom edges."""
    be = backend
    img = synth(40, 72, 11)
    img[:, 36:] = (img[:, 36:] // 3) * 3                                # flatter right half: larger clusters
    pts = pixel_features(img)
    d_pts = torch.from_numpy(pts).to(be.device)
    for eps, mp in ((1.0, 1), (1.5, 2), (2.0, 4), (2.9, 3), (3.0, 8), (4.0, 12), (4.9, 20)):
        plan = D.LatticeDbscan(be, 40, 72, eps, mp)
        lab, core = plan.run(d_pts)
        want = O.dbscan_labels(pts, eps, mp)
        assert np.array_equal(lab.cpu().numpy().astype(np.int64), want), (eps, mp)
    # anything that is not pixel (x, y) with 8-bit colours must be refused, wherever it sits in a tile
    plan = D.LatticeDbscan(be, 40, 72, 2.0, 3)
    for idx, col, val in ((0, 2, 0.5), (71, 3, 256.0), (72 * 20 + 67, 4, -1.0), (72 * 39 + 71, 0, 70.0),
                          (72 * 33 + 5, 1, 32.5), (72 * 32 + 64, 2, float("nan"))):
        bad = pts.copy()
        bad[idx, col] = val
        with pytest.raises(Exception):
            plan.run(torch.from_numpy(bad).to(be.device))


@pytest.mark.gpu
def test_lattice_equals_generic_on_1080p():
    from roibasedimagecompression_b200._lib import lib
    be = lib()
    img = synth(1080, 1920, 4321)
    pts = torch.from_numpy(pixel_features(img)).cuda()
    for eps, mp in ((3.0, 8), (2.0, 3), (5.0, 20)):
        lab, core = D.dbscan_image(be, torch.from_numpy(img.copy()).cuda(), eps, mp)
        lab_g, core_g = D.dbscan_points(be, pts, eps, mp)
        assert torch.equal(lab.view(-1), lab_g) and torch.equal(core.view(-1), core_g), (eps, mp)
