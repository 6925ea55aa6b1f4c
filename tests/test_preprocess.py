"""frame::Lidar::process_frame (L/src/sensors/lidar/frame.cpp:101-193 with sort_clouds :28-51 and split_clouds :53-99) --
SURVEY section 8(f) N3, the host preprocessing in front of register_frame.

CPU:  the C oracle (oracle/limu_oracle_frame.c) against the reference's own compiled frame.cpp (oracle/_ref) and against the
      committed golden fixture generated from it (tests/golden/fixtures_frame.npz).
GPU:  limu_preprocess_frame / limu_odom_register_msg (csrc/preprocess.cu) against the oracle and the fixture, through the C ABI.

Parity bar: bit-exact records (x, y, z, intensity, curvature), timestamps, segment sizes and segment times whenever the
curvature keys are distinct. std::sort in the reference is unstable, so points with EQUAL curvature come out in an
unspecified order there; with ties the comparison is exact on the key sequence and as multisets inside every run of equal keys.
Constant-rotation path (no per-point offset time): curvature derives from atan2f, where glibc and the device differ in the
last float bit: |d curvature| <= 2e-4 ms, everything else exact.
"""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "fixtures_frame.npz")
CFG = dict(min_range=5.0, max_range=100.0, min_angle=0.0, max_angle=360.0, frame_rate=10.0, num_scan_lines=16, frame_split_num=1)


def synth_mod():
    import __graft_entry__ as g
    g.load_package()
    from importlib import import_module
    return import_module("limu_b200.synth")


def make_msg(seed, n, beams=16, kind="distinct", message_time=1000.25, fields=None, point_step=None):
    """One PointCloud2 payload. kind: distinct (all offset times differ) | ties (columns share a firing time) |
    nooffset (timestamps all zero -> constant rotation model)."""
    synth = synth_mod()
    rng = np.random.default_rng(seed)
    xyz = (rng.normal(size=(n, 3)) * 30).astype(np.float32)
    xyz[rng.random(n) < 0.02] *= 0.05                     # inside the blind zone
    xyz[rng.random(n) < 0.02] *= 8.0                      # beyond max_range
    xyz[rng.random(n) < 0.01, rng.integers(0, 3)] = np.nan
    ring = rng.integers(0, beams, n)
    if kind == "distinct":
        stamp = message_time - 0.1 + rng.permutation(n) * (0.1 / n) + 1e-7
    elif kind == "ties":
        stamp = message_time - 0.1 + (rng.permutation(n) // beams) * (0.1 * beams / n) + 1e-7
        stamp[np.argmin(stamp)] -= 1e-4                   # a unique earliest point (the one split_clouds drops)
    else:
        stamp = np.zeros(n)
    kw = {}
    if fields is not None:
        kw = dict(fields=fields, point_step=point_step)
    data, fl = synth.make_pointcloud2(xyz, ring, stamp, intensity=rng.integers(0, 255, n), **kw)
    return data, fl, message_time


def assert_segments_equal(a, b, ties=False, curv_atol=0.0):
    assert len(a) == len(b)
    for sa, sb in zip(a, b):
        pa, pb = sa["points"], sb["points"]
        assert pa.shape == pb.shape
        assert sa["time"] == sb["time"]
        if curv_atol:
            # match rows by their coordinates (unique), compare curvature within tolerance, both sorted by their own key
            ka = np.lexsort((pa[:, 2], pa[:, 1], pa[:, 0])); kb = np.lexsort((pb[:, 2], pb[:, 1], pb[:, 0]))
            assert np.array_equal(pa[ka][:, :4], pb[kb][:, :4])
            assert np.abs(pa[ka][:, 4] - pb[kb][:, 4]).max() <= curv_atol
            assert np.all(np.diff(pa[:, 4]) >= 0) and np.all(np.diff(pb[:, 4]) >= 0)
            assert np.array_equal(sa["ts"][ka], sb["ts"][kb])
            continue
        assert np.array_equal(pa[:, 4], pb[:, 4])                   # the key sequence is always identical
        if ties:
            oa = np.lexsort((sa["ts"], pa[:, 3], pa[:, 2], pa[:, 1], pa[:, 0], pa[:, 4]))
            ob = np.lexsort((sb["ts"], pb[:, 3], pb[:, 2], pb[:, 1], pb[:, 0], pb[:, 4]))
            assert np.array_equal(pa[oa], pb[ob]) and np.array_equal(sa["ts"][oa], sb["ts"][ob])
        else:
            assert np.array_equal(pa, pb) and np.array_equal(sa["ts"], sb["ts"])


CASES = [  # (seed, n, kind, frame_split_num, scan_count)
    (1, 5000, "distinct", 1, 1), (2, 5000, "distinct", 3, 25), (3, 4999, "distinct", 4, 30), (4, 5000, "distinct", 4, 7),
    (5, 37, "distinct", 5, 30), (6, 3, "distinct", 1, 1), (7, 2, "distinct", 1, 1), (8, 1, "distinct", 1, 1),
    (9, 40, "distinct", 40, 50), (10, 2048, "ties", 1, 3), (11, 20000, "distinct", 2, 21),
]
NOOFFSET = [(21, 5000, 1, 1), (22, 5000, 2, 40), (23, 700, 3, 33)]


# ---------------------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("seed,n,kind,split,sc", CASES)
def test_port_matches_compiled_reference(ref, port, seed, n, kind, split, sc):
    data, fields, mt = make_msg(seed, n, kind=kind)
    cfg = dict(CFG, frame_split_num=split)
    a, b = ref.process_frame(data, fields, cfg, mt, sc), port.process_frame(data, fields, cfg, mt, sc)
    assert_segments_equal(a, b, ties=(kind == "ties"))
    if n >= 100:
        assert len(a) == (1 if sc < 20 else split) and sum(len(s["points"]) for s in a) > 0.8 * n


@pytest.mark.parametrize("seed,n,split,sc", NOOFFSET)
def test_port_matches_reference_constant_rotation_model(ref, port, seed, n, split, sc):
    data, fields, mt = make_msg(seed, n, kind="nooffset")
    cfg = dict(CFG, frame_split_num=split)
    a, b = ref.process_frame(data, fields, cfg, mt, sc), port.process_frame(data, fields, cfg, mt, sc)
    assert_segments_equal(a, b, ties=True)     # same libm on both sides: exact; many points share curvature 0..period


def test_port_timestamp_field_rules(ref, port):
    """utils::get_time_stamps: last of t/timestamp/time wins; 'time' is read as double and not normalised; none -> exception."""
    synth = synth_mod()
    F = synth
    base = [("x", 0, F.PF_FLOAT32, 1), ("y", 4, F.PF_FLOAT32, 1), ("z", 8, F.PF_FLOAT32, 1), ("intensity", 12, F.PF_UINT8, 1), ("ring", 14, F.PF_UINT16, 1)]
    # (a) a uint32 't' field in front of the double 'timestamp': 'timestamp' still wins for get_time_stamps (it comes last)
    fa = base + [("t", 24, F.PF_UINT32, 1), ("timestamp", 16, F.PF_FLOAT64, 1)]
    # (b) double 'time' last: used as is
    fb = base + [("timestamp", 16, F.PF_FLOAT64, 1), ("time", 24, F.PF_FLOAT64, 1)]
    # (c) no timestamp-like field at all
    for fields, expect_err in ((fa, False), (fb, False), (base, True)):
        data, fl, mt = make_msg(31, 3000, fields=fields, point_step=32)
        rng = np.random.default_rng(5)
        if any(f[0] == "t" for f in fields):
            data[:, 24:28] = rng.integers(0, 2 ** 32, len(data), dtype=np.uint32).reshape(-1, 1).view(np.uint8)
        if any(f[0] == "time" for f in fields):
            data[:, 24:32] = (rng.random(len(data)) * 3.0 - 1.0).reshape(-1, 1).view(np.uint8)
        a, b = ref.process_frame(data, fl, CFG, mt, 1), port.process_frame(data, fl, CFG, mt, 1)
        if expect_err:
            assert a == -1 and b == -1
        else:
            assert_segments_equal(a, b)


def test_port_matches_golden(port):
    g = np.load(GOLD)
    for k in range(int(g["n_cases"])):
        data, mt, sc, split = g[f"c{k}_data"], float(g[f"c{k}_mt"]), int(g[f"c{k}_sc"]), int(g[f"c{k}_split"])
        synth = synth_mod()
        out = port.process_frame(data, synth.LIDAR_POINT_FIELDS, dict(CFG, frame_split_num=split), mt, sc)
        sizes = g[f"c{k}_sizes"]
        assert [len(s["points"]) for s in out] == sizes.tolist()
        pts, ts = np.concatenate([s["points"] for s in out]), np.concatenate([s["ts"] for s in out])
        atol = 2e-4 if bool(g[f"c{k}_nooffset"]) else 0.0
        assert np.array_equal(pts[:, :4], g[f"c{k}_points"][:, :4]) and np.abs(pts[:, 4] - g[f"c{k}_points"][:, 4]).max() <= atol
        assert np.array_equal(ts, g[f"c{k}_ts"]) and np.array_equal([s["time"] for s in out], g[f"c{k}_times"])


def test_field_selection_rules_of_the_library():
    """limu_cloud_fields_from_pointfields is host code: it must apply pcl::fromROSMsg's name+datatype rule and get_time_stamps' rule."""
    import __graft_entry__ as g
    pkg = g.load_package()
    synth = synth_mod()
    cf = pkg.cloud_fields(synth.LIDAR_POINT_FIELDS, synth.LIDAR_POINT_STEP)
    assert (cf.off_x, cf.off_y, cf.off_z, cf.off_intensity, cf.off_ring, cf.off_timestamp, cf.off_time_field, cf.time_field_is_f64) == (0, 4, 8, 12, 14, 16, 16, 0)
    # a float intensity does not match the uint8 member; 'time' after 'timestamp' takes over the timestamp extraction
    f2 = [("x", 0, synth.PF_FLOAT32, 1), ("y", 4, synth.PF_FLOAT32, 1), ("z", 8, synth.PF_FLOAT32, 1), ("intensity", 12, synth.PF_FLOAT32, 1),
          ("timestamp", 16, synth.PF_FLOAT64, 1), ("time", 24, synth.PF_FLOAT64, 1)]
    cf = pkg.cloud_fields(f2, 32)
    assert (cf.off_intensity, cf.off_ring, cf.off_timestamp, cf.off_time_field, cf.time_field_is_f64) == (-1, -1, 16, 24, 1)
    with pytest.raises(pkg.LimuError, match="not existing"):
        pkg.cloud_fields(f2[:4], 32)


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    pkg = g.load_package()
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,kind,split,sc", CASES + [(12, 131072, "distinct", 3, 40)])
def test_cuda_matches_oracle(ctx, port, seed, n, kind, split, sc):
    data, fields, mt = make_msg(seed, n, kind=kind)
    cfg = dict(CFG, frame_split_num=split)
    a, b = port.process_frame(data, fields, cfg, mt, sc), ctx.process_frame(data, fields, cfg, mt, sc)
    assert_segments_equal(a, b)                 # oracle and device both sort stably: exact even with ties
    for s in b:                                 # the 48-byte PCL rows: data[3] = 1, normals and padding 0
        r = s["records"]
        assert np.all(r[:, 3] == 1.0) and not r[:, [4, 5, 6, 7, 10, 11]].any()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,split,sc", NOOFFSET)
def test_cuda_constant_rotation_model(ctx, port, seed, n, split, sc):
    data, fields, mt = make_msg(seed, n, kind="nooffset")
    cfg = dict(CFG, frame_split_num=split)
    a, b = port.process_frame(data, fields, cfg, mt, sc), ctx.process_frame(data, fields, cfg, mt, sc)
    if split == 1 or sc < 20:
        assert_segments_equal(a, b, curv_atol=2e-4)
    else:   # cut positions fall inside runs of (nearly) equal keys: compare the message as a whole
        assert [len(s["points"]) for s in a] == [len(s["points"]) for s in b]
        pa, pb = np.concatenate([s["points"] for s in a]), np.concatenate([s["points"] for s in b])
        ka = np.lexsort((pa[:, 2], pa[:, 1], pa[:, 0])); kb = np.lexsort((pb[:, 2], pb[:, 1], pb[:, 0]))
        assert np.array_equal(pa[ka][:, :4], pb[kb][:, :4])


@pytest.mark.gpu
def test_cuda_matches_golden(ctx):
    g = np.load(GOLD)
    synth = synth_mod()
    for k in range(int(g["n_cases"])):
        if bool(g[f"c{k}_nooffset"]):
            continue
        out = ctx.process_frame(g[f"c{k}_data"], synth.LIDAR_POINT_FIELDS, dict(CFG, frame_split_num=int(g[f"c{k}_split"])), float(g[f"c{k}_mt"]), int(g[f"c{k}_sc"]))
        assert [len(s["points"]) for s in out] == g[f"c{k}_sizes"].tolist()
        assert np.array_equal(np.concatenate([s["points"] for s in out]), g[f"c{k}_points"])
        assert np.array_equal(np.concatenate([s["ts"] for s in out]), g[f"c{k}_ts"])
        assert np.array_equal([s["time"] for s in out], g[f"c{k}_times"])


@pytest.mark.gpu
def test_cuda_errors_and_edges(ctx):
    import __graft_entry__ as g
    pkg = g.load_package()
    data, fields, mt = make_msg(41, 4000, kind="nooffset", beams=32)     # ring ids up to 31 against num_scan_lines = 16
    with pytest.raises(pkg.LimuError):
        ctx.process_frame(data, fields, CFG, mt, 1)
    data, fields, mt = make_msg(42, 4000)
    assert len(ctx.process_frame(data, fields, CFG, mt, 1)) == 1         # the status word was cleared
    assert ctx.process_frame(data[:0], fields, CFG, mt, 1) == []
    with pytest.raises(pkg.LimuError):
        ctx.process_frame(data, fields, dict(CFG, frame_split_num=0), mt, 1)
    far = data.copy()
    far[:, 0:4] = np.full((len(far), 1), 1.0e4, np.float32).view(np.uint8)   # everything beyond max_range
    assert ctx.process_frame(far, fields, CFG, mt, 1) == []


@pytest.mark.gpu
def test_full_size_properties(ctx):
    """BASELINE-size message (128k points): size-independent properties instead of an oracle run."""
    synth = synth_mod()
    scene = synth.Scene(seed=3)
    traj = synth.loop_trajectory(2, radius=30.0, step=1.0)
    scan = synth.pad_scan(synth.cast_scan(scene, traj[0], traj[1], beams=64, azimuth_steps=2000, seed=9), 128000)
    n = len(scan)
    rng = np.random.default_rng(0)
    perm = rng.permutation(n)                                            # the message arrives in arbitrary order
    mt = 500.0
    stamp = mt - 0.1 + (np.arange(n) + 0.5) * (0.1 / n)
    data, fields = synth.make_pointcloud2(scan[perm, :3], np.arange(n)[perm] % 64, stamp[perm])
    out = ctx.process_frame(data, fields, dict(CFG, num_scan_lines=64, frame_split_num=4), mt, 100)
    assert len(out) == 4
    d2 = (scan[:, 0] * scan[:, 0] + scan[:, 1] * scan[:, 1]) + scan[:, 2] * scan[:, 2]
    kept = np.flatnonzero((d2.astype(np.float64) >= 25.0) & (d2.astype(np.float64) <= 1.0e4))
    pts = np.concatenate([s["points"] for s in out])
    assert len(pts) == len(kept) - 1                                     # split_clouds never emits the first sorted point
    assert np.array_equal(pts[:, :3], scan[kept[1:], :3])                # sorted back into firing order
    for s in out:
        assert np.all(np.diff(s["points"][:, 4]) >= 0) and s["ts"].min() >= 0.0 and s["ts"].max() <= 1.0
    assert out[0]["time"] == mt and all(out[k]["time"] > out[k - 1]["time"] for k in range(1, 4))


@pytest.mark.gpu
def test_register_msg_matches_oracle_pipeline(ctx, port):
    """lidar_callback -> estimate_lidar_odometry on the device == oracle process_frame + oracle register_frame."""
    synth = synth_mod()
    scene = synth.Scene(seed=5)
    traj = synth.loop_trajectory(9, radius=30.0, step=0.3)
    cfg = dict(CFG, num_scan_lines=16, frame_split_num=1)
    ko = port.Kiss(voxel_size=1.0, max_range=100.0, cap=10, deskew=True, icp_max_iteration=100)
    kg = ctx.KissICP(voxel_size=1.0, max_range=100.0, cap=10, deskew=True, icp_max_iteration=100)
    rng = np.random.default_rng(8)
    for i in range(8):
        scan = synth.cast_scan(scene, traj[i], traj[i + 1], beams=16, azimuth_steps=500, seed=70 + i)
        n = len(scan)
        mt = 100.0 + 0.1 * i
        perm = rng.permutation(n)
        stamp = mt - 0.1 + (np.arange(n) + 0.5) * (0.1 / n)
        data, fields = synth.make_pointcloud2(scan[perm, :3], np.arange(n)[perm] % 16, stamp[perm])
        seg = port.process_frame(data, fields, cfg, mt, i + 1)
        assert len(seg) == 1
        _, _, pose_o = ko.register_cloud(seg[0]["points"][:, :3], seg[0]["ts"])
        poses, sizes, times, stats = kg.register_msg(data, fields, cfg, mt, i + 1)
        assert len(poses) == 1 and sizes[0] == len(seg[0]["points"]) and times[0] == seg[0]["time"]
        assert np.abs(poses[0][4:] - pose_o[4:]).max() < 1e-5 and np.abs(poses[0][:4] - pose_o[:4]).max() < 1e-6
        assert stats[0].n_points == sizes[0] and stats[0].icp.iterations == port._kiss_last_iterations(ko.h)
    kg.close()


def _random_seeds_e():
    """4 seeds in the suite; LIMU_RANDOM_SEEDS_E="lo-hi" widens the campaign (profiles/r2_random_campaign.json)."""
    spec = os.environ.get("LIMU_RANDOM_SEEDS_E", "")
    if "-" in spec:
        lo, hi = spec.split("-")
        return range(int(lo), int(hi))
    return range(4000, 4004)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", _random_seeds_e())
def test_cuda_matches_oracle_at_random_shapes(ctx, port, seed):
    """frame::Lidar::process_frame (frame.cpp:28-193) on the device against the oracle at seeded random shapes: message sizes 1 .. 150 000
    points (log-uniform), 1 .. 6 frame segments, any scan count (the first 20 messages are not split), distinct stamps or columns that
    share a firing time."""
    rng = np.random.default_rng(seed)
    n = int(np.exp(rng.uniform(0.0, np.log(150000.0))))
    ties = bool(rng.random() < 0.25)
    split = 1 if ties else int(rng.integers(1, 7))
    sc = int(rng.choice([1, 7, 19, 20, 21, 50]))
    beams = int(rng.choice([16, 32, 64]))
    data, fields, mt = make_msg(seed, n, beams=beams, kind="ties" if ties else "distinct")
    cfg = dict(CFG, frame_split_num=split, num_scan_lines=beams)
    a, b = port.process_frame(data, fields, cfg, mt, sc), ctx.process_frame(data, fields, cfg, mt, sc)
    assert_segments_equal(a, b)
