"""Kernel-mode measurement of the fused registration kernel (SURVEY section 8d C3/C5 shapes): every point of a large scan is an ICP query
against a ~45 M-point voxel map (>> L2), fixed iteration count. Same record as bench.py's `roofline_kernel_mode`.
   [LIMU_LIB=build/liblimu_<variant>.so] python tools/kernel_mode_bench.py [--queries 524288,4194304]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--queries", default="524288,4194304")
ap.add_argument("--voxels", type=float, default=2.5e6)
ap.add_argument("--fill", type=float, default=20.0)
ap.add_argument("--cap", type=int, default=20)
ap.add_argument("--voxel", type=float, default=0.5)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
pkg = g.load_package()
ctx = pkg.Context(0)
rec = bench.kernel_mode_record(torch, pkg, ctx, queries=tuple(int(x) for x in args.queries.split(",")), voxels=args.voxels, fill=args.fill, voxel=args.voxel,
                               cap=args.cap, iters=args.iters)
rec["library"] = os.environ.get("LIMU_LIB", "default")
print(json.dumps(rec))
