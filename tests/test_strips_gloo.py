"""Strip-sharded DBSCAN (eps halo + all-gather of boundary union-find edges) on world_size-2 and -3 gloo
groups on the CPU, kernels from the host-emulation build: every rank's labels must equal the unsplit run's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, EMU_SO, golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, H, W, img_seed, eps, mp_, out_dir, lattice=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from roibasedimagecompression_b200._lib import Backend
    from roibasedimagecompression_b200 import dbscan as D
    from roibasedimagecompression_b200.synth import synth, pixel_features
    be = Backend(EMU_SO, "cpu")
    img = synth(H, W, img_seed)
    pts = pixel_features(img)
    r0, r1, l0, l1, zone = D.strip_rows(H, world, rank, eps)
    if lattice:
        labels, core = D.dbscan_image_strips(be, torch.from_numpy(img[l0:l1].copy()), W, l0, (r0 - l0, r1 - l0),
                                             [(a - l0, b - l0) for a, b in zone], eps, mp_)
    else:
        local = torch.from_numpy(pts[l0 * W:l1 * W].copy())
        zone_idx = [((a - l0) * W, (b - l0) * W) for a, b in zone]
        labels, core = D.dbscan_strips(be, local, l0 * W, ((r0 - l0) * W, (r1 - l0) * W), zone_idx, eps, mp_)
    np.save(os.path.join(out_dir, f"lab{rank}.npy"), labels.numpy())
    np.save(os.path.join(out_dir, f"core{rank}.npy"), core.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,eps,mp_,lattice", [(2, 3.0, 4, False), (3, 2.0, 1, False), (2, 5.0, 8, False),
                                                  (2, 3.0, 4, True), (3, 5.0, 8, True)])
def test_strips_equal_unsplit(tmp_path, emu_backend, world, eps, mp_, lattice):
    H, W, seed = 48, 64, 1234                       # the image of tests/golden/dbscan_points.npz
    g = golden("dbscan_points.npz")
    combos = [tuple(c) for c in g["combos"]]
    from oracle import rhccq_oracle as O
    want = g[f"labels{combos.index((eps, float(mp_)))}"] if (eps, float(mp_)) in combos else \
        O.dbscan_labels(g["points"], eps, mp_)
    mp.spawn(_worker, args=(world, _free_port(), H, W, seed, eps, mp_, str(tmp_path), lattice), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"lab{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)


def test_strip_rows_cover_and_halo():
    from roibasedimagecompression_b200 import dbscan as D
    for H, world, eps in ((48, 2, 3.0), (100, 8, 2.5), (7, 3, 1.0)):
        rows = [D.strip_rows(H, world, r, eps) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == H
        for a, b in zip(rows[:-1], rows[1:]):
            assert a[1] == b[0]
        hz = int(np.ceil(2 * eps))
        for r0, r1, l0, l1, zone in rows:
            assert l0 == max(0, r0 - hz) and l1 == min(H, r1 + hz)
            for a, b in zone:
                assert l0 <= a < b <= l1
