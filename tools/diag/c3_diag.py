"""Diagnostic: configs[2] pipeline sequence (512k pts/scan, voxel 0.5, cap 20) with / without a resident background slab, plain and
pipelined path, against the C port scan by scan."""
import copy, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
import bench, oracle
pkg = g.load_package()
ctx = pkg.Context(0)
a3 = copy.copy(bench.parse([]))
a3.points, a3.beams, a3.azimuth_steps, a3.voxel, a3.cap, a3.max_range = 512000, 128, 4000, 0.5, 20, 1000.0
n = int(os.environ.get("N", 5))
nbg = int(os.environ.get("NBG", 6_000_000))
half = float(os.environ.get("HALF", 75.0))
z0 = float(os.environ.get("Z0", 150.0))
mr = float(os.environ.get("MAXR", 1000.0))
scans = bench.make_scans(a3, n, 42, "cuda:0", workload="c3")
port = oracle.load_port()
k = port.Kiss(voxel_size=0.5, max_range=mr, cap=20, deskew=True, icp_max_iteration=500)
rows = []
for s in scans:
    d, sr, p = k.register_cloud(np.ascontiguousarray(s[:, :3]), s[:, 3].astype(np.float64))
    rows.append((len(d), len(sr), k.last_iterations(), p.copy()))
staged = [torch.from_numpy(s).cuda() for s in scans]
gen = torch.Generator(device="cuda").manual_seed(3)
bg = torch.empty((nbg, 3), dtype=torch.float64, device="cuda")
bg[:, :2] = (torch.rand((nbg, 2), generator=gen, device="cuda", dtype=torch.float64) - 0.5) * 2 * half
bg[:, 2] = z0 + torch.rand(nbg, generator=gen, device="cuda", dtype=torch.float64) * 2.0
torch.cuda.synchronize()
for spec in (False, True):
    for with_bg in (False, True):
        o = ctx.KissICP(voxel_size=0.5, max_range=mr, cap=20, deskew=True, icp_max_iteration=500, speculate=spec, map_capacity_voxels=int(os.environ.get("CAPV", 600000)))
        print(f"--- speculate={spec} background={with_bg}")
        for i, t in enumerate(staged):
            if spec and i > 0 and i + 1 < len(staged):
                o.hint_next_dev(staged[i + 1].data_ptr(), 512000)
            t0 = time.perf_counter()
            p = o.register_frame_dev(t.data_ptr(), 512000)
            ctx.sync()
            dt = time.perf_counter() - t0
            st = o.stats
            r = rows[i]
            print(f"scan {i}: {dt*1e3:9.3f} ms  n_down {st.n_down} ({r[0]})  n_key {st.n_keypoints} ({r[1]})  iters {st.icp.iterations} ({r[2]})  kbar {st.icp.mean_candidates:.2f} fmiss {st.icp.miss_fraction:.3f} ncorr {st.icp.last_ncorr} "
                  f"dpose {np.abs(p - r[3]).max():.3e}")
            if i == 0 and with_bg:
                m = o.local_map()
                for lo in range(0, nbg, 1 << 20):
                    m.insert_points_dev(bg[lo:lo + (1 << 20)].data_ptr(), min(1 << 20, nbg - lo))
                ctx.sync()
                print("   map after background:", m.size())
        print("   map at end:", o.local_map().size())
        o.close()
