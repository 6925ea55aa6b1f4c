"""Developer tool: per-phase latency of one Gauss-Newton iteration inside k_icp_persistent.
Build the instrumented library first (not shipped):
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
       -DLIMU_ICP_PHASE_TIMING -shared -o lidar-imu-slam_b200/build/liblimu_pt.so lidar-imu-slam_b200/csrc/*.cu -lcudart_static
  LIMU_LIB=lidar-imu-slam_b200/build/liblimu_pt.so python tools/icp_phase_timing.py
With LIMU_ICP_PHASE_TIMING the hg_trace rows carry %globaltimer stamps of CTA 0 / thread 0:
0 loop top, 1 queries done, 2 CTA row written, 3 grid barrier passed, 4 rows folded, 5 solve done."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
ctx = pkg.Context(0)
rng = np.random.default_rng(0)
world = rng.random((300000, 3)) * np.array([100, 100, 6])
gm = ctx.VoxelHashMap(1.0, 100.0, 10)
gm.insert_points(world)
for nq in (2300, 20000, 128000, 1000000):
    src = world[rng.choice(len(world), nq, replace=nq > len(world))] + rng.normal(size=(nq, 3)) * 0.3
    init = np.array([0, 0, 0, 1.0, 0.05, 0.02, 0])
    for rep in range(3):
        r = gm.icp(src, init, 6.0, 2 / 3, 12, 1e-12, trace=True)
    t = r["hg"][:, :6]
    d = np.diff(t, axis=1)[1:]          # skip the first iteration
    names = ["queries", "cta_reduce", "grid_barrier", "fold", "solve"]
    print(f"nq={nq:8d} iters={r['iters']} per-iteration us:", {n: round(float(v) / 1e3, 2) for n, v in zip(names, d.mean(axis=0))},
          "total", round(float(np.diff(t[:, 0]).mean()) / 1e3, 2), "kbar", round(r["mean_candidates"], 2), "fmiss", round(r["miss_fraction"], 3))
