"""configs[2] in pipeline mode on the GPU alone (no CPU arms): the sequence of bench.py's `workload_c3` record -- 512 000 points per scan, voxel
0.5 m, cap 20, ~45 M-point background slab inserted behind scan 0 -- through the pipelined path, for profiling the kernels at that size with
ncu (tools/gpu_call_k.sh). Prints the stage split of the timed scans."""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
ctx = pkg.Context(0)
a3 = copy.copy(bench.parse([]))
a3.points, a3.beams, a3.azimuth_steps, a3.voxel, a3.cap, a3.max_range, a3.icp_mode = 512000, 128, 4000, 0.5, 20, 1000.0, 0
W3, K3 = 3, int(os.environ.get("K3", "6"))
bg_points = int(os.environ.get("NBG", "46000000"))
scans = bench.make_scans(a3, W3 + K3, 42, "cuda:0", workload="c3")
devs = [torch.from_numpy(s).cuda() for s in scans]
gen = torch.Generator(device="cuda").manual_seed(3)
bg = torch.empty((bg_points, 3), dtype=torch.float64, device="cuda")
bg[:, :2] = (torch.rand((bg_points, 2), generator=gen, device="cuda", dtype=torch.float64) - 0.5) * 400.0
bg[:, 2] = 150.0 + torch.rand(bg_points, generator=gen, device="cuda", dtype=torch.float64) * 2.0
torch.cuda.synchronize()
o = ctx.KissICP(voxel_size=a3.voxel, max_range=a3.max_range, cap=a3.cap, deskew=True, icp_max_iteration=a3.max_iter, icp_mode=0, speculate=True,
                map_capacity_voxels=3_400_000)
rows = []
for i, d in enumerate(devs):
    if i == W3:
        ctx.sync()
        ctx.set_profiling(True)
    if i > 0 and i + 1 < len(devs):
        o.hint_next_dev(devs[i + 1].data_ptr(), a3.points)
    o.register_frame_dev(d.data_ptr(), a3.points)
    st = o.stats
    rows.append((st.n_keypoints, st.icp.iterations, st.n_down))
    if i == 0:
        m = o.local_map()
        for lo in range(0, bg_points, 1 << 20):
            m.insert_points_dev(bg[lo:lo + (1 << 20)].data_ptr(), min(1 << 20, bg_points - lo))
ctx.sync()
prof, nfr = ctx.profile()
ctx.set_profiling(False)
nv, npts = o.local_map().size()
print(json.dumps({"scans": len(devs), "timed": nfr, "stage_ms_per_scan": {k: v / max(nfr, 1) for k, v in prof.items()}, "keypoints_iterations_ndown": rows,
                  "map_voxels": int(nv), "map_points": int(npts), "source_hash": pkg.source_hash()}))
o.close()
