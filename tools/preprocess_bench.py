"""Timing of frame::Lidar::process_frame on one BASELINE-size message (128 000 points, 64 rings):
the device path (limu_preprocess_frame through the host-pointer C ABI, H2D + D2H inside the timed region; and
limu_odom_register_msg = preprocess + register_frame from device memory) next to the reference's own frame.cpp on the host
(oracle/_ref) and the C oracle. Prints one JSON line.   python tools/preprocess_bench.py [--points 128000] [--reps 20]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=128000)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--split", type=int, default=1)
args = ap.parse_args()

pkg = g.load_package()
from importlib import import_module
synth = import_module("limu_b200.synth")
import oracle

scene = synth.Scene(seed=42)
traj = synth.loop_trajectory(2, radius=30.0, step=1.0)
scan = synth.pad_scan(synth.cast_scan(scene, traj[0], traj[1], beams=64, azimuth_steps=2000, seed=1), args.points)
n = len(scan)
rng = np.random.default_rng(0)
perm = rng.permutation(n)
mt = 500.0
stamp = mt - 0.1 + (np.arange(n) + 0.5) * (0.1 / n)
data, fields = synth.make_pointcloud2(scan[perm, :3], np.arange(n)[perm] % 64, stamp[perm])
cfg = dict(min_range=5.0, max_range=100.0, min_angle=0.0, max_angle=360.0, frame_rate=10.0, num_scan_lines=64, frame_split_num=args.split)

ctx = pkg.Context(0)
cf = pkg.cloud_fields(fields, data.shape[1])
lc = pkg.lidar_config(**cfg)
for _ in range(3):
    out = ctx.process_frame(data, cf, lc, mt, 100)
l0 = pkg.kernel_launches()
t0 = time.perf_counter()
for _ in range(args.reps):
    out = ctx.process_frame(data, cf, lc, mt, 100)
gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
launches = (pkg.kernel_launches() - l0) / args.reps

k = ctx.KissICP(voxel_size=1.0, max_range=100.0, cap=10, deskew=True)
for i in range(3):
    k.register_msg(data, cf, lc, mt, 100 + i)
t0 = time.perf_counter()
for i in range(args.reps):
    k.register_msg(data, cf, lc, mt, 103 + i)
msg_ms = (time.perf_counter() - t0) / args.reps * 1e3

line = {"tool": "preprocess_bench", "points": n, "kept": int(sum(len(s["points"]) for s in out)), "frame_split_num": args.split,
        "gpu_process_frame_ms": gpu_ms, "gpu_kernel_launches": launches, "gpu_register_msg_ms": msg_ms,
        "bytes_h2d": int(data.nbytes), "bytes_d2h": int(sum(s["records"].nbytes + s["ts"].nbytes for s in out))}
for name, api in (("reference", oracle.load_ref() if oracle.have_ref() else None), ("port", oracle.load_port())):
    if api is None:
        continue
    api.process_frame(data, fields, cfg, mt, 100)
    t0 = time.perf_counter()
    reps = max(3, args.reps // 4)
    for _ in range(reps):
        api.process_frame(data, fields, cfg, mt, 100)
    line[f"cpu_{name}_ms"] = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps(line))
