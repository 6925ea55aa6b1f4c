#!/usr/bin/env python
"""bench.py -- LiDAR odometry hot path throughput (BASELINE.json metric) on B200.

Workload (BASELINE.json configs[1], SURVEY section 8d "C2"): synthetic 64-beam spinning LiDAR, 128 000
points per scan, loop trajectory of radius 30 m at 1 m/scan, voxel 1.0 m, max_points_per_voxel 10,
deskew on. One STEP = one scan through KissICP::register_frame (deskew -> 0.5v/1.5v first-wins
downsample -> IQR -> fused correspondence + normal-equation + Gauss-Newton loop -> map insert/evict).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the B200 path (liblimu_cuda.so)
  python bench.py --impl reference ...                           the reference's own CPU code (oracle/_ref)

N > 1 (torchrun): one independent sequence per GPU (configs[3], "fleet replay"): weak scaling, no
data-path collective; torch.distributed is used only for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deskew+ICP+GN-loop odometry throughput at 128k pts/scan"
UNIT = "scans/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150, help="scans timed; the loop of radius 30 m closes after ~188 scans at 1 m/scan, where the reference's "
                    "Gauss-Newton loop stops converging and runs to its 500-iteration cap (both arms alike): W + K <= 180 stays in the converging regime")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=128000)
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--azimuth-steps", type=int, default=2000)
    ap.add_argument("--step-m", type=float, default=1.0, help="metres travelled per scan (10 m/s at 10 Hz)")
    ap.add_argument("--voxel", type=float, default=1.0)
    ap.add_argument("--cap", type=int, default=10)
    ap.add_argument("--max-iter", type=int, default=500)
    ap.add_argument("--icp-mode", type=int, default=0, help="0 = the reference's registration rules (the parity configuration, default); 1 = nearest of the "
                    "27-cell neighbourhood, 2 = point-to-plane, 3 = both (opt-in variants, SURVEY 8f N2; the CPU arm is then the C oracle)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="budget of the --impl reference run")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_scans(args, n_scans, seed, device):
    """The sequence of one vehicle: n_scans sweeps along the loop, each resized to exactly --points rows."""
    import __graft_entry__ as g
    g.load_package()
    from importlib import import_module
    synth = import_module("limu_b200.synth")
    scene = synth.Scene(seed=seed)
    traj = synth.loop_trajectory(n_scans + 1, radius=30.0, step=args.step_m)
    scans = []
    for i in range(n_scans):
        s = synth.cast_scan(scene, traj[i], traj[i + 1], beams=args.beams, azimuth_steps=args.azimuth_steps, seed=seed * 100003 + i, device=device)
        scans.append(synth.pad_scan(s, args.points, seed=i))
    return scans


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu --set full
    capture of this same command (profiles/r1_frame_kernels_ncu.json; cold cache). None if the capture is absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_frame_kernels_ncu.json")))
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for l in d["launches"]:
            if "k_icp_persistent" in l["kernel"]:
                r, ru = l["dram__bytes_read.sum"].split()
                w, wu = l["dram__bytes_write.sum"].split()
                vals.append(float(r) * unit[ru] + float(w) * unit[wu])
        return float(np.mean(vals)) if vals else None
    except Exception:
        return None


def k4_bytes(n_q, iters, kbar, fmiss):
    """Algorithmic bytes of the fused registration loop: per query per iteration
    24 B query + 16 B own-voxel slot + 24*kbar B candidate points + fmiss * 27 * 16 B fallback probes."""
    return iters * n_q * (24.0 + 16.0 + 24.0 * kbar + fmiss * 27 * 16.0)


def frame_kernel_bytes(n_q, iters, kbar, fmiss, n_src0, n_down, map_slots):
    """The persistent frame kernel also runs the IQR filter (24 B per candidate keypoint), the map insert (64 B per
    inserted point) and the eviction sweep (16 B per table slot) -- SURVEY section 8d K6 / K3."""
    return k4_bytes(n_q, iters, kbar, fmiss) + 24.0 * n_src0 + 64.0 * n_down + 16.0 * map_slots


def cpu_reference_api(icp_mode=0):
    import oracle
    if icp_mode == 0 and os.path.exists(oracle.REF_MT_SO):
        return oracle.load_ref(mt=True), "reference"
    return oracle.load_port(), "port"   # the opt-in variants do not exist in the reference: the C oracle defines them


def time_cpu(args, scans, warmup, max_steps, budget_s):
    """The reference's own register_frame on the host cores over a bounded prefix of the same sequence."""
    api, kind = cpu_reference_api(args.icp_mode)
    k = api.Kiss(voxel_size=args.voxel, max_range=100.0, cap=args.cap, deskew=True, icp_max_iteration=args.max_iter)
    if args.icp_mode:
        k.set_mode(args.icp_mode)
    xyz = [np.ascontiguousarray(s[:, :3]) for s in scans]
    ts = [s[:, 3].astype(np.float64) for s in scans]
    for i in range(min(warmup, len(scans))):
        k.register_cloud(xyz[i], ts[i])
    done, t0 = 0, time.perf_counter()
    for i in range(warmup, min(len(scans), warmup + max_steps)):
        k.register_cloud(xyz[i], ts[i])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done, dt, kind, api.num_threads()


def main():
    args = parse()
    rank, local_rank, world = dist_env()
    n_gpus = max(args.gpus, world)
    W, K = args.warmup, args.steps
    import torch

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        est_steps = max(1, min(K, 2000))
        scans = make_scans(args, W + est_steps, 42, dev)
        done, dt, kind, cores = time_cpu(args, scans, W, est_steps, args.ref_seconds)
        val = done / dt
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": n_gpus, "steps": done, "steps_requested": K, "warmup": W,
            "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[1]: synthetic {args.beams}-beam LiDAR, {args.points} pts/scan, loop r=30 m at {args.step_m} m/scan, voxel {args.voxel} m, cap {args.cap}, deskew on",
                       "icp_mode": args.icp_mode,
                       "note": "the reference's unmodified register_frame (oracle/_ref, thread-pool TBB shim) on the host cores; rank 0 only"
                               if kind == "reference" else "C oracle (single thread): the opt-in registration variant has no counterpart in the reference"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{done} consecutive scans after {W} warm-up scans (budget {args.ref_seconds:.0f} s)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mpoints_per_s": val * args.points / 1e6,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- the B200 path
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as g
    pkg = g.load_package()
    ctx = pkg.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    scans = make_scans(args, W + K, 42 + rank, f"cuda:{local_rank}")     # configs[3]: seeds 42..49, one sequence per GPU
    n_pts = args.points
    scan_bytes = n_pts * 16
    pinned = [pkg.PinnedArray((n_pts, 4), np.float32) for _ in scans]
    for p, s in zip(pinned, scans):
        p.array[...] = s
    dev_scans = [torch.from_numpy(s).cuda() for s in scans]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    flush.fill_(1)                                                      # push the staged scans out of L2
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()

    def new_odom():
        return ctx.KissICP(voxel_size=args.voxel, max_range=100.0, cap=args.cap, deskew=True, icp_max_iteration=args.max_iter, icp_mode=args.icp_mode)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)

    # -- arm 1: end to end through the host-pointer API (H2D of the scan + D2H of pose and both clouds every step)
    d2h = [0]

    def call_host(o, _s, i):
        if i + 1 < len(pinned):
            o.prefetch(pinned[i + 1].array)      # replay mode: the next scan's H2D overlaps this scan's kernels (same bytes, same step)
        down, key, _pose = o.register_frame(pinned[i].array, want_clouds=True, copy=False)
        d2h[0] += 56 + 16 + down.nbytes + key.nbytes

    odom = new_odom()
    for i in range(W):
        call_host(odom, None, i)
    barrier()
    d2h[0] = 0
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    t0 = time.perf_counter()
    for i in range(W, W + K):
        call_host(odom, None, i)
    e1.record(ext)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_s = max_over_ranks(max(e2e_wall, e0.elapsed_time(e1) / 1e3))
    e2e_d2h = d2h[0] / K
    odom.close()

    # -- arm 2: inputs resident in HBM (the `value`), with per-stage device timing for the roofline
    odom = new_odom()
    for i in range(W):
        odom.register_frame_dev(dev_scans[i].data_ptr(), n_pts)
    flush.fill_(2)
    barrier()
    ctx.set_profiling(True)
    launches0 = pkg.kernel_launches()
    frames = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    t0 = time.perf_counter()
    for i in range(W, W + K):
        if i + 1 < len(dev_scans):
            odom.hint_next_dev(dev_scans[i + 1].data_ptr(), n_pts)   # replay hint; a no-op unless the library was built with LIMU_SPECULATIVE_VOXELIZE
        odom.register_frame_dev(dev_scans[i].data_ptr(), n_pts)
        st = odom.stats
        frames.append((st.n_keypoints, st.icp.iterations, st.icp.mean_candidates, st.icp.miss_fraction, st.n_down))
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) / 1e3
    total_s = max_over_ranks(max(wall, dev_s))
    launches = pkg.kernel_launches() - launches0
    clocks = sampler.stop()
    prof, nframes = ctx.profile()
    ctx.set_profiling(False)
    last_pose = odom.poses()[-1]
    odom.close()

    fr = np.array(frames, dtype=np.float64)
    iters_total = float(fr[:, 1].sum())
    map_slots = 1 << 20   # C2 table: 2^20 slots (limu_odom_create: capacity 320k voxels -> next pow2 of 2x)
    alg_bytes = float(sum(frame_kernel_bytes(nk, it, kb, fm, nk * 1.05, nd, map_slots) for nk, it, kb, fm, nd in frames))
    icp_only_bytes = float(sum(k4_bytes(nk, it, kb, fm) for nk, it, kb, fm, _ in frames))
    icp_ms = prof["icp"]
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (icp_ms * 1e-3) / 1e9 if icp_ms > 0 else 0.0
    stage_share = {k: round(v / max(sum(prof.values()), 1e-9), 4) for k, v in prof.items()}

    cpu = None
    if rank == 0 and n_gpus == 1:
        done, dt, kind, cores = time_cpu(args, scans, W, K, args.cpu_seconds)
        cpu = {"value": done / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {done} timed scans of the same sequence after {W} warm-up scans ({dt:.1f} s of host time)"}

    if rank == 0:
        value = n_gpus * K / total_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * total_s / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"configs[1]: synthetic {args.beams}-beam LiDAR, {n_pts} pts/scan, loop r=30 m at {args.step_m} m/scan, voxel {args.voxel} m, cap {args.cap}, deskew on"
                            + (f"; configs[3]: {n_gpus} independent sequences, one per GPU" if n_gpus > 1 else ""),
                "l2": "L2 flushed (512 MB write) after staging; every step reads a different scan, none re-read; the local map is persistent state",
                "timing": "K steps bracketed by stream sync (+ barrier); CUDA events on the library stream and host wall clock, the larger one, max over ranks",
                "icp_max_iteration": args.max_iter,
                "icp_mode": args.icp_mode,
            },
            "mpoints_per_s": value * n_pts / 1e6,
            "iterations_per_scan": iters_total / K, "scans_at_iteration_cap": int((fr[:, 1] >= args.max_iter).sum()),
            "keypoints_per_scan": float(fr[:, 0].mean()), "downsampled_per_scan": float(fr[:, 4].mean()),
            "k_bar": float(fr[:, 2].mean()), "f_miss": float(fr[:, 3].mean()),
            "e2e": {"value": n_gpus * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": scan_bytes, "d2h_bytes_per_step": int(e2e_d2h),
                    "api": "limu_odom_register_frame (host pointers, pinned) with limu_odom_prefetch of the following scan: every step uploads one 2 MB scan (overlapped with the previous step's kernels) and reads back pose + downsampled + keypoint clouds"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_icp_persistent<latency> (one launch per scan: IQR filter + fused correspondence/residual/Jacobian/normal-equation Gauss-Newton loop + map insert + eviction sweep)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": profiled_traffic(),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / max(nframes, 1), "avg_launch_ms": icp_ms / max(nframes, 1),
                         "share_of_step": stage_share.get("icp"),
                         "registration_loop_bytes_per_launch": icp_only_bytes / max(nframes, 1),
                         "note": "pipeline mode: ~2.3k keypoint queries per iteration -> latency bound by construction (SURVEY H3): each Gauss-Newton iteration is a grid-wide dependency chain of ~10 us; the HBM-bound shape of the same kernel (4M queries vs a 45M-point map, 51% of measured HBM peak) is in profiles/ (tools/kernel_mode_bench.py)"},
            "stage_ms_per_step": {k: v / max(nframes, 1) for k, v in prof.items()}, "stage_share": stage_share,
            "clocks": clocks,
            "last_pose": [round(float(x), 6) for x in last_pose],
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
