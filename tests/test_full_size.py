"""The hot path at BASELINE.json's FULL sizes (configs[1]: 128 000 points/scan; configs[2]: 512 000 points/scan, voxel 0.5 m,
cap 20), checked through size-independent properties and independent numpy restatements of the integer/index work instead of
a slow oracle run:

  * voxel keys and first-point-wins downsampling   == numpy trunc + first occurrence per key (exact)
  * IQR keypoint filter                            == numpy sort + Tukey fence (exact)
  * capped ordered map insert                      == "first `cap` points of every voxel in input order", voxels in creation order (exact)
  * neighbour search                               idempotence: every stored point is its own nearest neighbour at distance 0
  * rigid transform                                round trip T^-1 (T p) == p to 1e-9
  * normal equations at 512k queries               linearity: H/g of the whole query set == H/g of its halves added (1e-12 relative)
  * whole pipeline at 128k points                  determinism: two runs and the device-pointer entry give bit-identical poses;
                                                   counts equal the numpy restatement of voxelize
"""
import numpy as np
import pytest

from np_restatement import np_first_per_voxel, np_iqr, np_keys, np_map_insert   # pinned against the C oracle by tests/test_np_restatement.py

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def synth(pkg):
    from importlib import import_module
    return import_module("limu_b200.synth")


def big_scan(synth, beams, az, n, seed, street=False):
    scene = synth.Scene(seed=seed, street=street)
    traj = synth.loop_trajectory(2, radius=30.0, step=1.0)
    return synth.pad_scan(synth.cast_scan(scene, traj[0], traj[1], beams=beams, azimuth_steps=az, seed=seed), n, seed=seed)


@pytest.mark.parametrize("n,beams,az,vox", [(128000, 64, 2000, 1.0), (512000, 128, 4000, 0.5)])
def test_keys_downsample_iqr_at_full_size(ctx, synth, n, beams, az, vox):
    scan = big_scan(synth, beams, az, n, seed=11, street=(n > 200000))
    xyz = scan[:, :3].astype(np.float64)
    assert np.array_equal(ctx.voxel_keys(xyz, vox), np_keys(xyz, vox).astype(np.int32))
    first = np_first_per_voxel(xyz, 0.5 * vox)
    down, idx = ctx.voxel_downsample(xyz, 0.5 * vox, with_index=True)
    assert np.array_equal(idx, first) and np.array_equal(down, xyz[first])
    first2 = np_first_per_voxel(down, 1.5 * vox)
    src0 = ctx.voxel_downsample(down, 1.5 * vox)
    assert np.array_equal(src0, down[first2])
    assert np.array_equal(ctx.iqr_processing(src0), np_iqr(src0))
    assert np.array_equal(ctx.iqr_processing(xyz), np_iqr(xyz))               # the filter itself at full size
    src, down2 = ctx.voxelize(xyz, vox)                                        # the fused kernel against the chain above
    assert np.array_equal(down2, down) and np.array_equal(src, np_iqr(src0))


def test_map_insert_neighbours_transform_at_full_size(ctx, synth):
    cap, vox = 20, 0.5
    scan = big_scan(synth, 128, 4000, 512000, seed=12, street=True)
    rng = np.random.default_rng(3)
    xyz = np.concatenate([scan[:, :3].astype(np.float64) + rng.normal(size=(len(scan), 3)) * 0.05 for _ in range(3)])   # 1.5 M points
    m = ctx.VoxelHashMap(vox, 1.0e4, cap)
    for part in np.array_split(xyz, 3):
        m.insert_points(part)
    keys, counts, pts = m.dump()
    # voxels in order of first occurrence, each holding its first `cap` points in input order (three batches == one batch)
    ek, ec, ep = np_map_insert(xyz, vox, cap)
    assert np.array_equal(keys, ek) and np.array_equal(counts, ec) and np.array_equal(pts, ep)
    assert m.size() == (len(ek), len(ep))
    # idempotence: a stored point is its own nearest neighbour (its voxel exists and holds it at distance 0)
    probe = pts[:: max(1, len(pts) // 600000)]
    got, gkey, grank = m.get_closest_neighbour(probe, with_index=True)
    assert np.array_equal(got, probe) and (grank >= 0).all()
    assert np.array_equal(gkey, np_keys(probe, vox).astype(np.int32))
    # correspondences within tau of themselves: every probe pairs with itself
    s_, t_ = m.get_correspondences(probe, 0.3)
    assert len(s_) == len(probe) and np.array_equal(s_, t_)
    # rigid transform round trip
    T = np.array([0.0, 0.0, np.sin(0.35), np.cos(0.35), 12.5, -3.25, 0.75])
    Ti = np.empty(7)
    import ctypes as C
    from importlib import import_module
    lib = import_module("limu_b200").lib()
    lib.limu_se3_inverse(T.ctypes.data_as(C.POINTER(C.c_double)), Ti.ctypes.data_as(C.POINTER(C.c_double)))
    back = ctx.transform_points(Ti, ctx.transform_points(T, probe))
    assert np.abs(back - probe).max() < 1e-9
    m.close()


def test_normal_equations_are_additive_at_full_size(ctx, synth):
    """lidar::align_clouds' J^T J / J^T r reduction over 512k queries (kernel mode, the bandwidth shape): the sums over the whole set
    equal the sums over its halves added -- a checksum of checksums that does not need a CPU pass over 512k queries."""
    vox, cap = 0.5, 20
    scan = big_scan(synth, 128, 4000, 512000, seed=13, street=True)
    world = scan[:, :3].astype(np.float64)
    m = ctx.VoxelHashMap(vox, 1.0e4, cap)
    m.insert_points(world)
    init = np.array([0.0, 0.0, np.sin(0.001), np.cos(0.001), 0.05, -0.03, 0.01])
    q = world + np.random.default_rng(5).normal(size=world.shape) * 0.02
    full = m.icp(q, init, 1.5, 0.5, 1, 1e-12, trace=True)
    a = m.icp(q[: len(q) // 2], init, 1.5, 0.5, 1, 1e-12, trace=True)
    b = m.icp(q[len(q) // 2:], init, 1.5, 0.5, 1, 1e-12, trace=True)
    assert full["ncorr"][0] == a["ncorr"][0] + b["ncorr"][0] > 0.9 * len(q)
    scale = np.abs(full["hg"]).max()
    assert np.abs(full["hg"][0] - (a["hg"][0] + b["hg"][0])).max() < 1e-12 * scale
    # and the solve of the summed system is the estimate the full run produced
    assert full["iters"] == 1
    m.close()


def test_pipeline_at_128k_is_deterministic_and_consistent(ctx, pkg, synth):
    import torch
    scene = synth.Scene(seed=42)
    traj = synth.loop_trajectory(6, radius=30.0, step=1.0)
    scans = [synth.pad_scan(synth.cast_scan(scene, traj[i], traj[i + 1], beams=64, azimuth_steps=2000, seed=42 * 100003 + i), 128000, seed=i) for i in range(5)]
    runs = []
    for kind in ("host", "host", "dev", "dev+hint"):
        k = ctx.KissICP(voxel_size=1.0, max_range=100.0, cap=10, deskew=True)
        poses = []
        for s in scans:
            if kind == "host":
                down, src, pose = k.register_frame(s)
                xyz = s[:, :3].astype(np.float64)
                if k.stats.deskewed == 0:                                     # raw scan -> numpy restatement of voxelize applies directly
                    f1 = np_first_per_voxel(xyz, 0.5)
                    f2 = np_first_per_voxel(xyz[f1], 1.5)
                    assert np.array_equal(down, xyz[f1]) and np.array_equal(src, np_iqr(xyz[f1][f2]))
                assert k.stats.n_points == 128000 and k.stats.n_down == len(down) and k.stats.n_keypoints == len(src)
            else:
                if kind == "dev":
                    t = torch.from_numpy(s).cuda()
                    torch.cuda.synchronize()                                  # the library runs on its own non-blocking stream
                else:   # replay hints (limu_odom_hint_next_dev): results must not depend on them, whatever the build does with them
                    if not poses:
                        staged = [torch.from_numpy(x).cuda() for x in scans]
                        torch.cuda.synchronize()
                    t = staged[len(poses)]
                    if len(poses) + 1 < len(staged):
                        k.hint_next_dev(staged[len(poses) + 1].data_ptr(), len(scans[len(poses) + 1]))
                pose = k.register_frame_dev(t.data_ptr(), len(s))
            poses.append(pose.copy())
        runs.append(np.array(poses))
        nv, npts = k.local_map().size()
        assert nv > 1000 and npts >= nv
        k.close()
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    # a build that acts on the hints computes the deskew twist on the device: equal to the north-star tolerance, bit-equal otherwise
    assert np.abs(runs[3] - runs[0]).max() < 1e-9
