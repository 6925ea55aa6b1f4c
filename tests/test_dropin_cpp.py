"""The C++ drop-in layer (include/limu_dropin: lidar::VoxelHashMap, align_clouds, ICP, MotionCompensator, KissICP with
the reference's own signatures) driven from a real C++ program (tests/cpp/dropin_main.cpp, prebuilt where the reference
headers exist) and checked against the oracle on the same inputs."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "tests", "cpp", "build", "dropin_test")


def write_arrays(path, arrs):
    with open(path, "wb") as f:
        f.write(struct.pack("<q", len(arrs)))
        for a in arrs:
            a = np.ascontiguousarray(a, np.float64).ravel()
            f.write(struct.pack("<q", a.size))
            f.write(a.tobytes())


def read_arrays(path):
    out = []
    with open(path, "rb") as f:
        (n,) = struct.unpack("<q", f.read(8))
        for _ in range(n):
            (m,) = struct.unpack("<q", f.read(8))
            out.append(np.frombuffer(f.read(8 * m), np.float64).copy())
    return out


def test_dropin_headers_name_the_reference_interfaces():
    """CPU check: every drop-in header exists and declares the reference's class / function names."""
    base = os.path.join(ROOT, "include", "limu_dropin", "limu", "sensors", "lidar")
    want = {"helpers/voxel_hash_map.hpp": ["class VoxelHashMap", "insert_points", "get_closest_neighbour", "get_correspondences",
                                           "remove_points_from_far", "pointcloud", "clear", "empty", "update"],
            "helpers/registration.hpp": ["align_clouds", "SE3d ICP("],
            "helpers/deskew.hpp": ["class MotionCompensator", "deskew_scan"],
            "frame.hpp": ["class Lidar", "struct ProcessingInfo", "initialize", "process_frame", "buffer_empty", "get_lidar_buffer_front",
                          "get_segment_ts_front", "curr_acc_segment_time", "pop", "set_current_pose_nav", "return_prev_ts"],
            "icp.hpp": ["class KissICP", "register_frame", "voxelize", "iqr_processing", "get_prediction_model", "get_adaptive_threshold",
                        "current_vel", "has_moved", "local_map_", "poses_"]}
    for rel, names in want.items():
        src = open(os.path.join(base, rel)).read()
        for n in names:
            assert n in src, (rel, n)


@pytest.mark.gpu
def test_cpp_consumer_matches_oracle(tmp_path, port, rng):
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/build/dropin_test not built (needs the reference headers at build time)")
    import __graft_entry__ as g
    g.load_package()
    from importlib import import_module
    synth = import_module("limu_b200.synth")
    A, B = rng.normal(size=(4000, 3)) * 10, rng.normal(size=(4000, 3)) * 10
    q = rng.normal(size=(3000, 3)) * 11
    tau, origin = 1.3, np.array([1.0, -2.0, 0.5])
    T = port.se3_exp(np.array([0.5, 0.2, -0.1, 0.01, 0.02, -0.05]))
    asrc = rng.normal(size=(2000, 3)) * 15
    atgt = port.transform(port.se3_exp(rng.normal(size=6) * 0.03), asrc) + rng.normal(size=(2000, 3)) * 0.01
    n = 30000
    pl = [rng.random((n // 3, 3)) * 40 - 20 for _ in range(3)]
    pl[0][:, 2] = rng.normal(size=n // 3) * 0.02
    pl[1][:, 0] = 20 + rng.normal(size=n // 3) * 0.02
    pl[2][:, 1] = -20 + rng.normal(size=n // 3) * 0.02
    world = np.concatenate(pl)
    true = port.se3_exp(np.array([0.3, -0.2, 0.05, 0.004, -0.003, 0.02]))
    isrc = port.transform(port.se3_inv(true), world[rng.choice(len(world), 2500, replace=False)])
    init = np.array([0, 0, 0, 1.0, 0, 0, 0])
    scene = synth.Scene(seed=5)
    traj = synth.loop_trajectory(7, radius=30.0, step=0.5)
    scans = [synth.cast_scan(scene, traj[i], traj[i + 1], beams=16, azimuth_steps=400, seed=40 + i) for i in range(6)]
    ts = [np.sort(rng.random(len(s))) for s in scans]          # genuine float64 timestamps (not float32-representable)
    arrs = [A, B, q, [tau], origin, T, asrc, atgt, [0.5], world, isrc, init, [6.0, 2.0 / 3.0, 60, 1e-4], [len(scans)]]
    for s, t in zip(scans, ts):
        arrs += [s[:, :3].astype(np.float64), t]
    # frame::Lidar drop-in: one PointCloud2 payload replayed 21 times (the 21st is split in three, frame.cpp:64)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_preprocess import CFG, make_msg
    msg, mfields, mt = make_msg(77, 3000)
    arrs += [msg.reshape(-1).astype(np.float64), [mt, msg.shape[1], 3, 21]]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_arrays(fin, arrs)
    r = subprocess.run([BIN, fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    o = read_arrays(fout)
    # oracle on the same inputs
    m = port.Map(1.0, 30.0, 5)
    assert o[0][0] == 1.0
    m.insert(A)
    m.insert(B)
    assert o[1][0] == m.num_voxels()
    assert np.array_equal(o[2].reshape(-1, 3), m.dump()[2])                       # pointcloud() in creation order
    assert np.array_equal(o[3].reshape(-1, 3), m.closest(q[:64]))
    s_, t_ = m.correspondences(q, tau)
    assert np.array_equal(o[4].reshape(-1, 3), s_) and np.array_equal(o[5].reshape(-1, 3), t_)
    m.remove_far(origin)
    assert np.array_equal(o[6].reshape(-1, 3), m.dump()[2])
    m.update(B, T)
    assert np.array_equal(o[7].reshape(-1, 3), m.dump()[2])
    np.testing.assert_allclose(o[8], port.align(asrc, atgt, 0.5)["pose"], rtol=0, atol=1e-10)
    w = port.Map(1.0, 100.0, 20)
    w.insert(world)
    np.testing.assert_allclose(o[9], port.icp(w, isrc, init, 6.0, 2.0 / 3.0, 60, 1e-4, trace=True)["pose"], rtol=0, atol=1e-9)
    k = port.Kiss(voxel_size=1.0, max_range=100.0, cap=10, deskew=True, icp_max_iteration=100)
    j = 10
    for i, (s, t) in enumerate(zip(scans, ts)):
        d, sr, p = k.register_cloud(s[:, :3], t)
        assert np.abs(o[j][4:] - p[4:]).max() < 1e-5 and np.abs(o[j][:4] - p[:4]).max() < 1e-6
        assert o[j + 1][0] == len(d) and o[j + 1][1] == len(sr)
        j += 2
        if i == len(scans) - 1:
            np.testing.assert_allclose(o[j].reshape(-1, 3), d, rtol=0, atol=1e-9)
            np.testing.assert_allclose(o[j + 1].reshape(-1, 3), sr, rtol=0, atol=1e-9)
            j += 2
    assert o[j][0] == len(scans)
    poses = k.poses()
    np.testing.assert_allclose(o[j + 1], port.se3_mul(port.se3_inv(poses[-2]), poses[-1]), rtol=0, atol=1e-6)
    dk = port.deskew(scans[-1][:, :3], ts[-1], poses[-2], poses[-1])
    np.testing.assert_allclose(o[j + 2].reshape(-1, 3), dk, rtol=0, atol=1e-5)
    seg = port.process_frame(msg, mfields, dict(CFG, frame_split_num=3), mt, 21)
    assert len(seg) == 3 and o[j + 3].tolist() == [len(s["points"]) for s in seg]
    assert np.array_equal(o[j + 4], [s["time"] for s in seg])
    assert np.array_equal(o[j + 5].reshape(-1, 5), np.concatenate([s["points"] for s in seg]).astype(np.float64))
    assert np.array_equal(o[j + 6], np.concatenate([s["ts"] for s in seg]))


def test_dropin_ekf_cpp_consumer_matches_python_binding():
    """include/limu_dropin/limu/kalman/ekf.hpp (kalman::EKF on the C ABI) from a real C++ program, against the same script through ctypes.
    Host code: no GPU needed."""
    exe = os.path.join(ROOT, "tests", "cpp", "build", "ekf_test")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/build/ekf_test not built (needs vendored Eigen at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = np.array([float(x) for x in out.stdout.split()])
    import __graft_entry__ as g
    pkg = g.load_package()
    e = pkg.Ekf(lidar_pose_trail=3)
    grav, trans = np.array([0, 0, -9.81]), np.array([0.05, -0.02, 0.1])
    rot = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    e.initialize_orientation(np.array([0.3, -0.2, 9.7]), grav)
    for i in range(20):
        e.predict(100.0 + 0.005 * i, np.array([0.02, -0.01, 0.2 + 0.001 * i]), np.array([0.3, 0.1 * i, 9.81]), grav, trans, rot)
    e.normalize_quaternions(True)
    e.update_and_propagate()
    e.update_lidar_pose(np.array([0.0, 0.0, 0.1, 0.99498743710662, 0.4, -0.1, 0.05]), 0.05, 0.01)
    m, P, t = e.state()
    want = np.concatenate([m, [np.trace(P), t, np.linalg.norm(m[3:6])]])
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-12
    e.close()
