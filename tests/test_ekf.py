"""kalman::EKF predict / update: the library's Eigen-free restatement (host code, csrc/ekf.cu) against the reference's
own ekf.cpp compiled into oracle/_ref (oracle/ref_ekf_driver.cpp drives its predict / zero_vel_update / update_visual_pose_aug /
update_undo_augmentation), and against golden states recorded from it (tests/golden/fixtures_ekf.npz) where the compiled reference is absent.
Tolerances: the state to 1e-12, the covariance to 1e-12 of its largest entry (dense products are summed in a different order than Eigen's
blocked GEMM; exp(S) is closed-form vs Eigen's Pade approximant). Host code: runs without a GPU."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures_ekf.npz")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


def imu_stream(rng, n, t0=100.0, rate=200.0):
    """A plausible IMU stream: slow yaw + pitch wobble, accelerations around gravity, in the reference's units."""
    t = t0 + np.arange(n) / rate
    xg = np.stack([0.02 * np.sin(0.7 * (t - t0)), 0.015 * np.cos(0.5 * (t - t0)), 0.2 + 0.05 * np.sin(0.3 * (t - t0))], 1) + rng.normal(size=(n, 3)) * 1e-3
    xa = np.stack([0.3 * np.sin(0.9 * (t - t0)), 0.2 * np.cos(0.4 * (t - t0)), np.full(n, 9.81)], 1) + rng.normal(size=(n, 3)) * 2e-2
    return t, xg, xa


def run_script(f, steps, grav, trans, rot):
    """Drive a filter (product or reference: same method names) through a list of steps, recording (m, P) after each."""
    out = []
    for s in steps:
        if s[0] == "predict":
            f.predict(s[1], s[2], s[3], grav, trans, rot)
        elif s[0] == "normalize":
            f.normalize_quaternions(s[1])
        elif s[0] == "zupt":
            f.zero_velocity_update(s[1])
        elif s[0] == "augment":
            f.augment_pose_trail()
        elif s[0] == "undo":
            f.undo_augmentation()
        m, P, _ = f.state()
        out.append((m.copy(), P.copy()))
    return out


def script(rng, n_imu=60):
    t, xg, xa = imu_stream(rng, n_imu)
    steps = []
    for i in range(n_imu):
        steps.append(("predict", t[i], xg[i], xa[i]))
        if i % 5 == 4:
            steps.append(("normalize", True))
        if i % 20 == 19:
            steps.append(("augment",))       # one scan = 20 IMU samples at 200 Hz / 10 Hz
        if i == 45:
            steps.append(("zupt", 1e-3))
        if i == 50:
            steps.append(("undo",))
    return steps


GRAV = np.array([0.0, 0.0, -9.81])
TRANS = np.array([0.05, -0.02, 0.10])
ROT = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])   # trace 1 > 0 branch of the matrix -> quaternion conversion


def compare(got, ref, m_tol=1e-12):
    assert len(got) == len(ref)
    for k, ((gm, gP), (rm, rP)) in enumerate(zip(got, ref)):
        assert np.abs(gm - rm).max() < m_tol, (k, np.abs(gm - rm).max())
        scale = max(np.abs(rP).max(), 1e-30)
        assert np.abs(gP - rP).max() < 1e-12 * scale, (k, np.abs(gP - rP).max() / scale)


@pytest.mark.parametrize("trail,noise_scale,ori_initialised", [(20, 1.0, True), (3, 2.5, True), (20, 1.0, False)])
def test_ekf_matches_compiled_reference(pkg, trail, noise_scale, ori_initialised):
    """ori_initialised: the orientation block of P as initialize_imu_global_orientation leaves it (diag(1,1,1,0) * init_ori_noise^2) -- the
    state then agrees to 1e-12. With the constructor's unit variances the augmentation's measurement update trusts the zero-initialised
    trail slot (variance 1e-6) a million times more than the orientation itself: it shrinks the quaternion by ~1e6 and renormalises it, so
    1e-16 rounding differences between two equally valid summation orders become ~1e-10 in the state (covariances still agree to 1e-15):
    that case is held to 1e-8."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/liblimu_ref.so not built (needs /root/reference at build time)")
    rng = np.random.default_rng(7)
    steps = script(rng)
    a = pkg.Ekf(lidar_pose_trail=trail, noise_scale=noise_scale)
    b = oracle.RefEkf(lidar_pose_trail=trail, noise_scale=noise_scale)
    assert a.dim == b.dim == 30 + 7 * trail
    m0, P0, _ = a.state()
    r0, Q0, _ = b.state()
    assert np.array_equal(m0, r0) and np.array_equal(P0, Q0)               # constructor: state and initial covariance bit for bit
    # start from a tilted, moving state with biases so that every Jacobian block is exercised
    m0[3:6] = [0.8, -0.3, 0.05]
    q = np.array([0.98, 0.05, -0.12, 0.1]); m0[6:10] = q / np.linalg.norm(q)
    m0[10:13] = [1e-3, -2e-3, 5e-4]; m0[13:16] = [0.02, -0.01, 0.03]; m0[16:19] = [1.01, 0.99, 1.02]; m0[19:22] = GRAV
    if ori_initialised:
        P0[6:10, 6:10] = np.diag([1e-6, 1e-6, 1e-6, 0.0]) * noise_scale ** 2
    a.set_state(m0, P0); b.set_state(m0, P0)
    tol = 1e-12 if ori_initialised else 1e-8
    compare(run_script(a, steps, GRAV, TRANS, ROT), run_script(b, steps, GRAV, TRANS, ROT), tol)
    # matrix -> quaternion: the other branch (trace <= 0), and dt <= 0 (skipped predict)
    rot2 = np.array([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 1.0]])
    t_last = steps[-1][1] if steps[-1][0] == "predict" else [s for s in steps if s[0] == "predict"][-1][1]
    more = [("predict", t_last + 0.005, np.array([0.1, 0.0, 0.3]), np.array([0.0, 0.1, 9.7]))]
    compare(run_script(a, more, GRAV, TRANS, rot2), run_script(b, more, GRAV, TRANS, rot2), tol)
    assert abs(a.state()[2] - b.state()[2]) < 1e-12                         # get_current_time
    # dt <= 0: the sample is skipped (ekf.cpp:235-240). Product only: the reference's message `std::cout << dt` crashes inside this container's
    # mixed libstdc++ (the oracle is built with /opt/gcc, Python loads the system runtime first); the branch has no arithmetic to compare.
    before = a.state()
    a.predict(t_last + 0.005, np.zeros(3), np.zeros(3), GRAV, TRANS, rot2)
    after = a.state()
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    a.close()


def test_ekf_matches_golden_states(pkg):
    """The same script against states recorded from the compiled reference (tests/golden/make_golden.py): no oracle/_ref needed."""
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/fixtures_ekf.npz missing")
    g = np.load(GOLD)
    rng = np.random.default_rng(int(g["seed"]))
    steps = script(rng)
    a = pkg.Ekf(lidar_pose_trail=int(g["trail"]), noise_scale=float(g["noise_scale"]))
    a.set_state(g["m0"], g["P0"])
    got = run_script(a, steps, GRAV, TRANS, ROT)
    idx = g["idx"]
    for k, i in enumerate(idx):
        assert np.abs(got[i][0] - g["m"][k]).max() < 1e-12
        scale = np.abs(g["P"][k]).max()
        assert np.abs(got[i][1] - g["P"][k]).max() < 1e-12 * scale
    a.close()


def test_update_and_propagate_and_orientation(pkg):
    """Glue (ekf.cpp:680-698) and the two functions without a usable original: initialize_imu_global_orientation (undefined behaviour in
    the reference: restated as intended) and the registration-pose measurement update (no counterpart), checked against closed forms."""
    a = pkg.Ekf(lidar_pose_trail=4)
    # stationary filter: zero-velocity update, newest pose dropped, then the augmentation -- no crash, symmetric covariance, unit quaternions
    a.predict(10.0, np.zeros(3), np.array([0.0, 0.0, 9.81]), GRAV, TRANS, ROT)
    a.predict(10.005, np.zeros(3), np.array([0.0, 0.0, 9.81]), GRAV, TRANS, ROT)
    m, P, _ = a.state()
    m[3:6] = [1e-5, 0, 0]
    a.set_state(m, None)
    a.update_and_propagate()
    m, P, t = a.state()
    assert np.abs(P - P.T).max() == 0.0 and abs(np.linalg.norm(m[6:10]) - 1) < 1e-12 and abs(t - 10.005) < 1e-12
    assert np.abs(m[30:33] - m[0:3]).max() < 1e-6 and np.abs(m[33:37] - m[6:10]).max() < 1e-6      # slot 0 of the trail cloned the current pose
    # orientation from gravity: R(q) maps calc_grav onto the accelerometer direction
    xa = np.array([0.3, -0.2, 9.7])
    a.initialize_orientation(xa, GRAV)
    m, P, _ = a.state()
    w, x, y, z = m[6:10]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    assert np.abs(R @ (GRAV / np.linalg.norm(GRAV)) - xa / np.linalg.norm(xa)).max() < 1e-12
    assert np.array_equal(m[19:22], GRAV) and P[9, 9] == 0.0 and P[6, 6] > 0
    # registration pose as a measurement: the textbook update on POS / ORI
    b = pkg.Ekf(lidar_pose_trail=2)
    m, P, _ = b.state()
    pose = np.array([0.02, -0.01, 0.05, 0.998, 1.5, -0.4, 0.2])
    pose[:4] /= np.linalg.norm(pose[:4])
    H = np.zeros((7, b.dim)); H[0:3, 0:3] = np.eye(3); H[3:7, 6:10] = np.eye(4)
    y = np.concatenate([pose[4:], [pose[3], pose[0], pose[1], pose[2]]])
    Rm = np.diag([0.05 ** 2] * 3 + [0.01 ** 2] * 4)
    S = H @ P @ H.T + Rm
    K = P @ H.T @ np.linalg.inv(S)
    m_exp = m + K @ (y - H @ m)
    P_exp = P - K @ H @ P
    m_exp[6:10] /= np.linalg.norm(m_exp[6:10]); m_exp[25:29] /= np.linalg.norm(m_exp[25:29])
    b.update_lidar_pose(pose, 0.05, 0.01)
    m2, P2, _ = b.state()
    assert np.abs(m2 - m_exp).max() < 1e-12 and np.abs(P2 - P_exp).max() < 1e-15
    a.close(); b.close()


def _random_seeds_k():
    """4 seeds in the suite; LIMU_RANDOM_SEEDS_K="lo-hi" widens the campaign (profiles/r2_random_campaign.json)."""
    import os
    spec = os.environ.get("LIMU_RANDOM_SEEDS_K", "")
    if "-" in spec:
        lo, hi = spec.split("-")
        return range(int(lo), int(hi))
    return range(5000, 5004)


@pytest.mark.parametrize("seed", _random_seeds_k())
def test_ekf_matches_compiled_reference_on_random_scripts(pkg, seed):
    """kalman::EKF (ekf.cpp:11-290, 521-764) behind limu_ekf_* against the compiled reference on seeded random scripts: IMU stream length,
    pose-trail length, noise scale, initial state, extrinsics, and WHERE the quaternion normalisation, the pose-trail augmentation, its
    undo and the zero-velocity update fall between the predictions. State 1e-11, covariance 1e-12 relative."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/liblimu_ref.so not built (needs /root/reference at build time)")
    rng = np.random.default_rng(seed)
    trail = int(rng.choice([3, 5, 10, 20]))
    noise_scale = float(rng.uniform(0.5, 3.0))
    n_imu = int(rng.integers(30, 110))
    t, xg, xa = imu_stream(rng, n_imu, t0=float(rng.uniform(0.0, 1e4)), rate=float(rng.choice([100.0, 200.0, 400.0])))
    steps, augmented = [], 0
    for i in range(n_imu):
        steps.append(("predict", t[i], xg[i], xa[i]))
        r = rng.random()
        if r < 0.15:
            steps.append(("normalize", bool(rng.integers(0, 2))))
        elif r < 0.25:
            steps.append(("augment",))
            augmented += 1
        elif r < 0.30 and augmented:
            steps.append(("undo",))
            augmented -= 1
        elif r < 0.35:
            steps.append(("zupt", float(10.0 ** rng.uniform(-4, -1))))
    a = pkg.Ekf(lidar_pose_trail=trail, noise_scale=noise_scale)
    b = oracle.RefEkf(lidar_pose_trail=trail, noise_scale=noise_scale)
    m0, P0, _ = a.state()
    m0[0:3] = rng.normal(size=3) * 5.0
    m0[3:6] = rng.normal(size=3)
    q = rng.normal(size=4); q[0] = abs(q[0]) + 0.5; m0[6:10] = q / np.linalg.norm(q)
    m0[10:13] = rng.normal(size=3) * 2e-3; m0[13:16] = rng.normal(size=3) * 2e-2; m0[16:19] = 1.0 + rng.normal(size=3) * 0.01; m0[19:22] = GRAV
    P0[6:10, 6:10] = np.diag([1e-6, 1e-6, 1e-6, 0.0]) * noise_scale ** 2
    a.set_state(m0, P0); b.set_state(m0, P0)
    trans = rng.normal(size=3) * 0.1
    rot = [ROT, np.eye(3), np.array([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, 1.0]])][int(rng.integers(0, 3))]
    compare(run_script(a, steps, GRAV, trans, rot), run_script(b, steps, GRAV, trans, rot), 1e-11)
    a.close()
