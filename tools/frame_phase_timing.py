"""Developer tool: where the time goes inside the fused frame kernel (IQR | ICP loop | insert claim | insert place |
eviction sweep) on the bench workload. Needs the instrumented build:
  LIMU_LIB=lidar-imu-slam_b200/build/liblimu_pt.so python tools/frame_phase_timing.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
from importlib import import_module
synth = import_module("limu_b200.synth")
ctx = pkg.Context(0)
scene = synth.Scene(seed=42)
K = 40
traj = synth.loop_trajectory(K + 1, radius=30.0, step=1.0)
scans = [synth.pad_scan(synth.cast_scan(scene, traj[i], traj[i + 1], seed=42 * 100003 + i, device="cuda"), 128000, seed=i) for i in range(K)]
odo = ctx.KissICP(voxel_size=1.0, cap=10, deskew=True, icp_max_iteration=500)
marks = np.zeros(16)
rows, iters = [], []
for i, s in enumerate(scans):
    odo.register_frame(s, want_clouds=False)
    pkg.lib().limu_debug_frame_marks(marks.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if i >= 5:
        rows.append(np.diff(marks[:6]) / 1e3)
        iters.append(odo.stats.icp.iterations)
r = np.array(rows)
names = ["iqr", "icp_loop", "insert_claim", "insert_place", "evict_sweep"]
print({n: round(float(v), 2) for n, v in zip(names, r.mean(axis=0))}, "total_us", round(float(r.sum(axis=1).mean()), 2), "iters/scan", np.mean(iters),
      "icp_us_per_iter", round(float(r[:, 1].mean() / np.mean(iters)), 2))
print("classic shape, solve breakdown (last iteration of the last scan, ns, 512 ns ticks): ldlt", marks[9] - marks[8], "exp", marks[10] - marks[9])
print("cluster shape, last iteration of the last scan (SM cycles, clock64): row reduce + DSMEM push + cluster barrier", marks[7] - marks[6], "fold", marks[8] - marks[7],
      "ldlt", marks[9] - marks[8], "exp", marks[10] - marks[9], "| pass of warp 1", marks[12] - marks[11], "tail of the solver warp", marks[14] - marks[13],
      "| raw marks 6..14 relative to mark 11:", [int(marks[k] - marks[11]) for k in (6, 7, 8, 9, 10, 11, 12, 13, 14)])
