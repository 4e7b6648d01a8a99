#!/usr/bin/env python
"""BASELINE.json configs[4]: the DBSCAN-only sweep — N in {1, 4, 16, 64} M pixel-feature points (1024^2 .. 8192^2
synthetic images, float32 [n,5]) x eps in {2, 3, 5, 8} x minPts in {1, 4, 8, 16} through the lattice engine, the
same through the generic cell-binning engine up to 16 M points, and uniform points in [0,256)^5 (true cell
binning).  Per combination: points/s of the whole DBSCAN (CUDA events, 3 runs after 2 warm-ups) and the neighbour
count's achieved algorithmic GB/s (24 B/point) against the measured HBM peak.

    python tools/c5_grid.py [--out profiles/r02_c5_grid.json] [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from roibasedimagecompression_b200 import dbscan as D  # noqa: E402
from roibasedimagecompression_b200._lib import lib  # noqa: E402
from roibasedimagecompression_b200.synth import synth, pixel_features  # noqa: E402

BYTES_PER_POINT = 24


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def image_points(side: int):
    t = min(side, 2048)
    img = np.concatenate([np.concatenate([synth(t, t, 4321 + r * (side // t) + c) for c in range(side // t)], axis=1)
                          for r in range(side // t)], axis=0)
    return torch.from_numpy(pixel_features(img)).cuda()


def time_plan(be, plan, pts, count_name, reps=3):
    for _ in range(2):
        labels, core = plan.run(pts)
    torch.cuda.synchronize()
    be.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        labels, core = plan.run(pts)
    e1.record(); torch.cuda.synchronize()
    kt = be.kernel_times_ms(); be.kernel_timing(False)
    ms = e0.elapsed_time(e1) / reps
    cn, cms = kt[count_name]
    n = pts.shape[0]
    ach = BYTES_PER_POINT * n / 1e9 / ((cms / cn) / 1e3)
    return {"ms": ms, "points_per_s": n / (ms / 1e3), "count_ms": cms / cn, "count_gbs": ach, "count_frac": ach / peak(),
            "clusters": int(labels.max().item()) + 1, "core_fraction": float(core.float().mean().item()),
            "phases_ms": {k: t / c for k, (c, t) in kt.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_c5_grid.json"))
    ap.add_argument("--quick", action="store_true", help="1 M and 16 M only, fewer minPts")
    a = ap.parse_args()
    be = lib()
    sides = [1024, 4096] if a.quick else [1024, 2048, 4096, 8192]
    eps_grid = [2.0, 3.0, 5.0, 8.0]
    mp_grid = [1, 8] if a.quick else [1, 4, 8, 16]
    rows = []
    t_start = time.time()
    for side in sides:
        pts = image_points(side)
        n = side * side
        for eps in eps_grid:
            for mp in mp_grid:
                plan = D.LatticeDbscan(be, side, side, eps, mp)
                r = time_plan(be, plan, pts, "rhccq_dbscan_lattice_count")
                rows.append(dict(engine="lattice", n=n, eps=eps, min_pts=mp, **r))
                del plan
                if n <= (1 << 24):
                    lo, hi = D.point_bounds(be, pts, 2)
                    plan = D.PointDbscan(be, n, 5, eps, mp, lo, hi, 2)
                    g = time_plan(be, plan, pts, "rhccq_dbscan_count")
                    rows.append(dict(engine="generic", n=n, eps=eps, min_pts=mp, labels_equal_lattice=None, **g))
                    del plan
                print(f"{n:>9d} pts eps {eps} minPts {mp:2d}: lattice {r['ms']:8.3f} ms, count {r['count_frac']:.3f} of HBM", flush=True)
        del pts
        torch.cuda.empty_cache()
    for n in ([1 << 22] if a.quick else [1 << 22, 1 << 24]):
        u = torch.from_numpy(np.random.default_rng(0).uniform(0, 256, size=(n, 5)).astype(np.float32)).cuda()
        for eps in eps_grid:
            for mp in ([1, 8] if a.quick else [1, 4, 8, 16]):
                lo, hi = D.point_bounds(be, u, 3)
                plan = D.PointDbscan(be, n, 5, eps, mp, lo, hi, 3)
                g = time_plan(be, plan, u, "rhccq_dbscan_count")
                rows.append(dict(engine="generic, uniform points in [0,256)^5, 3-D cells", n=n, eps=eps, min_pts=mp, **g))
                del plan
        del u
    out = {"command": "python tools/c5_grid.py" + (" --quick" if a.quick else ""), "hbm_peak_gbs": peak(),
           "bytes_per_point": BYTES_PER_POINT, "wall_s": time.time() - t_start, "rows": rows}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("wrote", a.out, len(rows), "rows")


if __name__ == "__main__":
    main()
