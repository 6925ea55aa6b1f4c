"""Point-sharded ICP over N GPUs (BASELINE configs[4] / SURVEY C5): correctness against the single-GPU run and timing
of the fused peer-mailbox exchange against the un-fused NCCL baseline. Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_icp_check.py
Prints one JSON line per case on rank 0; exit code 1 on any mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--queries", default="100000,4194304")
ap.add_argument("--voxels", type=float, default=1.0e6)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-nccl", action="store_true", help="skip the un-fused NCCL baseline")
ap.add_argument("--alarm", type=int, default=180, help="hard wall-clock limit of this process (s)")
args = ap.parse_args()
import signal
signal.alarm(args.alarm)      # never hold a multi-GPU box on a hang


def log(msg):
    print(f"[rank {os.environ.get('RANK', 0)}] {msg}", file=sys.stderr, flush=True)

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
# (Round 1 forced --no-nccl above 2 ranks because the baseline "hung in ncclAllReduce": it was k_icp_step spinning on a query
# stride of 0 -- IcpArgs::icp_blocks unset in icp_sharded_nccl -- and never reaching the all-reduce at ANY rank count.)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = g.load_package()
from importlib import import_module
sh = import_module("limu_b200.sharding")
ctx = pkg.Context(local)
log("context ok")
sh.connect(ctx, nccl_baseline=not args.no_nccl)
log("communicator connected")
ext = torch.cuda.ExternalStream(ctx.stream())

# replicated map: every rank inserts the same seeded points
voxel, cap = 0.5, 20
side = float(np.sqrt(args.voxels) * voxel)
gen = torch.Generator(device="cuda").manual_seed(1)
m = ctx.VoxelHashMap(voxel, 1e9, cap, capacity_voxels=int(args.voxels * 1.3))
total, done = int(args.voxels * 12), 0
while done < total:
    n = min(1 << 20, total - done)
    p = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    p[:, :2] = (torch.rand((n, 2), generator=gen, device="cuda", dtype=torch.float64) - 0.5) * side
    p[:, 2] = torch.randn(n, generator=gen, device="cuda", dtype=torch.float64) * 0.02 + 0.1
    torch.cuda.synchronize()
    m.insert_points_dev(p.data_ptr(), n)
    done += n
log("map built")
sizes = [None] * world
dist.all_gather_object(sizes, m.size())
ok = all(s == sizes[0] for s in sizes)

def timed(fn):
    best, out = None, None
    for rep in range(args.reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        out = fn()
        e1.record(ext)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep >= 1:
            best = float(t.item()) if best is None else min(best, float(t.item()))
    return best, out

init = pkg.se3_exp(np.array([0.03, -0.02, 0.01, 0.0005, -0.0003, 0.001]))
for nq in [int(x) for x in args.queries.split(",")]:
    gq = torch.Generator(device="cuda").manual_seed(7)
    q = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    q[:, :2] = (torch.rand((nq, 2), generator=gq, device="cuda", dtype=torch.float64) - 0.5) * side * 0.98
    q[:, 2] = torch.randn(nq, generator=gq, device="cuda", dtype=torch.float64) * 0.02 + 0.1
    torch.cuda.synchronize()
    lo, hi = sh.shard_range(nq, rank, world)
    shard = q[lo:hi].contiguous()
    single = m.icp_dev(q.data_ptr(), nq, init, 1.5, 0.5, args.iters, 1e-9)            # every rank: the 1-GPU answer
    log(f"nq={nq}: single-GPU done ({single['iters']} iterations)")
    t_fused, fused = timed(lambda: m.icp_sharded_dev(shard.data_ptr(), hi - lo, init, 1.5, 0.5, args.iters, 1e-9, mode=0))
    log(f"nq={nq}: fused sharded done")
    if args.no_nccl:
        t_nccl, nccl = None, fused   # not measured: never report the fused time under the baseline's name
    else:
        t_nccl, nccl = timed(lambda: m.icp_sharded_dev(shard.data_ptr(), hi - lo, init, 1.5, 0.5, args.iters, 1e-9, mode=1))
        log(f"nq={nq}: NCCL baseline done")
    t_single, _ = timed(lambda: m.icp_dev(q.data_ptr(), nq, init, 1.5, 0.5, args.iters, 1e-9))
    poses = [None] * world
    dist.all_gather_object(poses, (fused["pose"].tolist(), nccl["pose"].tolist(), fused["iters"], nccl["iters"]))
    same_across_ranks = all(p[0] == poses[0][0] for p in poses)                       # fused: bit-identical on every rank
    dt = np.abs(fused["pose"][4:] - single["pose"][4:]).max()
    dr = np.abs(fused["pose"][:4] - single["pose"][:4]).max()
    dtn = np.abs(nccl["pose"][4:] - single["pose"][4:]).max()
    drn = np.abs(nccl["pose"][:4] - single["pose"][:4]).max()
    good = (same_across_ranks and fused["iters"] == single["iters"] and nccl["iters"] == single["iters"] and fused["last_ncorr"] == single["last_ncorr"]
            and dt < 1e-9 and dr < 1e-10 and dtn < 1e-9 and drn < 1e-10)
    ok = ok and good
    if rank == 0:
        it = max(fused["iters"], 1)
        print(json.dumps({"ranks": world, "queries": nq, "iters": fused["iters"], "ncorr": fused["last_ncorr"], "ok": bool(good),
                          "pose_diff_vs_1gpu": {"fused_m": float(dt), "fused_q": float(dr), "nccl_m": float(dtn), "nccl_q": float(drn)},
                          "bit_identical_across_ranks": bool(same_across_ranks),
                          "us_per_iter": {"single_gpu": round(t_single * 1e3 / it, 2), "sharded_fused": round(t_fused * 1e3 / it, 2), "sharded_nccl": None if t_nccl is None else round(t_nccl * 1e3 / it, 2)},
                          "speedup_vs_single": round(t_single / t_fused, 3), "fused_vs_nccl": None if t_nccl is None else round(t_nccl / t_fused, 3)}))
ctx.comm_destroy()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
