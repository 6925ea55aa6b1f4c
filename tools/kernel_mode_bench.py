"""Kernel-mode measurement of the fused registration kernel (SURVEY section 8d C3/C5 shapes): every point of a large
scan is an ICP query against a voxel map much larger than L2, fixed iteration count. Prints one JSON line per case with
the algorithmic bytes (SURVEY formula with the measured k-bar / f_miss), the per-iteration time and the fraction of the
measured HBM peak.   python tools/kernel_mode_bench.py [--voxels 2.5e6] [--cap 20] [--queries 524288,4194304]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--voxel", type=float, default=0.5)
ap.add_argument("--cap", type=int, default=20)
ap.add_argument("--voxels", type=float, default=2.5e6, help="occupied voxels of the map (ground sheet)")
ap.add_argument("--fill", type=float, default=20.0, help="mean points offered per voxel")
ap.add_argument("--queries", default="524288,4194304")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

pkg = g.load_package()
ctx = pkg.Context(0)
side = float(np.sqrt(args.voxels) * args.voxel)          # square ground sheet, one voxel layer thick
gen = torch.Generator(device="cuda").manual_seed(1)
m = ctx.VoxelHashMap(args.voxel, 1e9, args.cap, capacity_voxels=int(args.voxels * 1.3))
total = int(args.voxels * args.fill)
t0 = time.time()
done = 0
while done < total:
    n = min(1 << 20, total - done)
    p = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    p[:, :2] = (torch.rand((n, 2), generator=gen, device="cuda", dtype=torch.float64) - 0.5) * side
    p[:, 2] = torch.randn(n, generator=gen, device="cuda", dtype=torch.float64) * 0.02 + 0.1
    torch.cuda.synchronize()
    m.insert_points_dev(p.data_ptr(), n)
    done += n
nv, npts = m.size()
print(f"# map: {nv} voxels, {npts} points ({npts * 24 / 1e9:.2f} GB of points), built in {time.time() - t0:.1f}s", file=sys.stderr)
peak = 6552.6
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
ext = torch.cuda.ExternalStream(ctx.stream())
for nq in [int(x) for x in args.queries.split(",")]:
    q = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    q[:, :2] = (torch.rand((nq, 2), generator=gen, device="cuda", dtype=torch.float64) - 0.5) * side * 0.98
    q[:, 2] = torch.randn(nq, generator=gen, device="cuda", dtype=torch.float64) * 0.02 + 0.1
    torch.cuda.synchronize()
    init = pkg.se3_exp(np.array([0.03, -0.02, 0.01, 0.0005, -0.0003, 0.001]))
    best = None
    for rep in range(args.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        r = m.icp_dev(q.data_ptr(), nq, init, 1.5, 0.5, args.iters, 0.0)
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rep >= 2:
            best = ms if best is None else min(best, ms)
    per_iter_us = best * 1e3 / r["iters"]
    kb, fm = r["mean_candidates"], r["miss_fraction"]
    bytes_iter = nq * (24 + 16 + 24 * kb + fm * 27 * 16)
    gbs = bytes_iter / (per_iter_us * 1e-6) / 1e9
    print(json.dumps({"queries": nq, "iters": r["iters"], "us_per_iter": round(per_iter_us, 2), "k_bar": round(kb, 3), "f_miss": round(fm, 4),
                      "alg_MB_per_iter": round(bytes_iter / 1e6, 1), "achieved_GBs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4),
                      "map_voxels": nv, "map_points": npts, "voxel": args.voxel, "cap": args.cap, "ncorr": r["last_ncorr"]}))
