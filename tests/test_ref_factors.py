"""Pins the oracle to the REFERENCE'S OWN source text (oracle/_ref/libref.so = /root/reference/src/factor/*.h,
src/utilies/common.h compiled unmodified against the stub Eigen / Ceres tree, tests/ref_lib.py).

Every comparison is oracle (oracle/*.hpp, the checker of all CUDA parity tests) against reference text on the same
inputs: residuals <= 1e-12 relative, Jet Jacobians <= 1e-11, on seeded random inputs, on corridor-like poses
(|q| ~ 2 rad, where the quaternion branches matter) and on the frozen tests/golden/factors.npz."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import lvio2d_b200 as L
import ref_lib
from lvio2d_b200.params import params_T

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not built and /root/reference absent")
RNG = np.random.default_rng(20261017)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rand_rotvec(max_angle=np.pi * 0.999):
    v = RNG.normal(size=3)
    return v / np.linalg.norm(v) * RNG.uniform(0, max_angle)


def corridor_like_pose():
    T_io = params_T(L.corridor_params(), "T_imu_to_wheel")
    Rwb = Rotation.from_euler("zyx", [RNG.uniform(-3, 3), RNG.normal(0, 2e-3), RNG.normal(0, 2e-3)]).as_matrix()
    return np.concatenate([RNG.uniform(-5, 5, 3), Rotation.from_matrix(Rwb @ T_io[:, :3].T).as_rotvec()])


def rand_pose():
    return np.concatenate([RNG.uniform(-5, 5, 3), rand_rotvec(3.0)])


def rand_state(pose=None):
    pose = corridor_like_pose() if pose is None else pose
    return np.concatenate([pose, RNG.normal(0, 0.5, 3), RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])


def test_reference_text_is_what_runs():
    """libref.so exports the driver's entry points and was linked from the reference's translation units."""
    import subprocess
    path = ref_lib.build()
    syms = subprocess.run(["nm", "-DC", path], capture_output=True, text=True, check=True).stdout
    for s in ("lvio_2d::solver::solve(", "lvio_2d::solver::marginalization(", "lvio_2d::laser_manager::do_match(",
              "lvio_2d::laser_manager::spawn_scan(", "convert::laser_to_point_times(", "marginalization_matrix("):
        assert s in syms, s


def test_lie_primitives(oracle):
    for _ in range(300):
        v = rand_rotvec()
        R = ref_lib.exp_so3(v)
        assert np.abs(R - oracle.exp_so3(v)).max() < 1e-15 + 2e-16
        assert np.abs(R - Rotation.from_rotvec(v).as_matrix()).max() < 1e-13
        assert np.abs(ref_lib.log_SO3(R) - oracle.log_SO3(R)).max() < 1e-14
    assert np.array_equal(ref_lib.exp_so3(np.zeros(3)), np.eye(3))
    for ang in (0.5, 3.0, np.pi + 0.1, 2 * np.pi + 0.4, 5 * np.pi - 0.2, 7.3):
        a = np.array([0.3, -0.2, 0.9])
        a = a / np.linalg.norm(a) * ang
        assert np.abs(ref_lib.normalize_so3(a) - oracle.normalize_so3(a)).max() < 1e-14   # 1 ulp of |a| ~ 15 (FMA contraction in the oracle build)
    for _ in range(100):
        th, d = rand_rotvec(), RNG.normal(0, 0.5, 3)
        out, jac = ref_lib.so3_plus(th, d, want_jacobian=True)
        assert np.abs(out - oracle.so3_plus(th, d)).max() < 1e-15
        assert np.abs(jac - np.eye(3)).max() < 1e-15   # |theta| <= pi: Plus is a plain sum at delta = 0
    for _ in range(100):
        p, p1, p2 = RNG.normal(size=3), RNG.normal(size=3), RNG.normal(size=3)
        assert abs(ref_lib.dis_from_line(p, p1, p2) - oracle.dis_from_line(p, p1, p2)) < 1e-14


def test_laser_factor(oracle, params):
    worst_r = worst_j = 0.0
    for k in range(60):
        pi_ = corridor_like_pose() if k % 2 else rand_pose()
        pj_ = pi_.copy()
        pj_[0:3] += RNG.normal(0, 0.3, 3)
        pj_[3:6] = (Rotation.from_rotvec(pi_[3:6]) * Rotation.from_rotvec(RNG.normal(0, 0.05, 3))).as_rotvec()
        l = [np.append(RNG.uniform(-6, 6, 2), 0) for _ in range(4)]
        r0, J0 = ref_lib.eval_laser_factor(*l, pi_, pj_)
        r1, J1 = oracle.eval_laser_factor(params, *l, pi_, pj_)
        worst_r, worst_j = max(worst_r, rel(r1, r0)), max(worst_j, rel(J1, J0))
        w = ref_lib.laser_pair_weight(*l)
        assert abs(w - np.sqrt(min(np.linalg.norm(l[0] - l[1]), np.linalg.norm(l[2] - l[3])) / 2.0 / 0.02)) < 1e-14
    assert worst_r < 1e-12 and worst_j < 1e-11, (worst_r, worst_j)


def make_imu_blob(oracle, params, n=20):
    samples = np.zeros((n, 7))
    samples[:, 0] = RNG.uniform(0.002, 0.006, n)
    samples[:, 1:4] = RNG.normal(0, 1.0, (n, 3)) + np.array([0.2, 9.7, 0.1])
    samples[:, 4:7] = RNG.normal(0, 0.3, (n, 3))
    bias = np.concatenate([RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])
    return samples, bias


def test_imu_preintegration(oracle, params):
    for n in (1, 5, 40):
        samples, bias = make_imu_blob(oracle, params, n)
        off = np.array([0, n], np.int64)
        b0 = ref_lib.imu_preintegrate(off, samples, bias[None])[0]
        b1 = oracle.imu_preintegrate(params, off, samples, bias[None])[0]
        assert rel(b1[0:15], b0[0:15]) < 1e-12          # X
        assert rel(b1[15:240], b0[15:240]) < 1e-12      # J
        assert rel(b1[240:465], b0[240:465]) < 1e-8     # sqrt_inverse_P: inverse + Cholesky of a matrix with cond ~ 1e8
        assert b1[465] == pytest.approx(b0[465], rel=1e-15)
        # the same interval through add_imu_measure / update_only_t with time stamps (what trajectory.cpp calls)
        stamps = np.concatenate([[100.0], 100.0 + np.cumsum(samples[:, 0])])
        b2 = ref_lib.imu_preintegrate_stamped(stamps, samples[:, 1:7], bias)
        assert rel(b2[0:240], b0[0:240]) < 1e-9          # dt = difference of stamps: ~1e-14 relative
    # the `last_ba` term of F(gamma, gamma) (imu_preintegraption.h:192) shows up when the accelerometer bias is large
    samples, bias = make_imu_blob(oracle, params, 10)
    bias[0:3] = [0.5, -0.4, 0.3]
    off = np.array([0, 10], np.int64)
    assert rel(oracle.imu_preintegrate(params, off, samples, bias[None])[0][15:240], ref_lib.imu_preintegrate(off, samples, bias[None])[0][15:240]) < 1e-12


def test_wheel_preintegration(oracle, params):
    for n in (1, 4, 25):
        steps = np.zeros((n, 7))
        steps[:, 0] = RNG.uniform(0.01, 0.05, n)
        steps[:, 1:4] = RNG.normal(0, 0.5, (n, 3)) * np.array([1, 0.2, 0.01])
        steps[:, 4:7] = RNG.normal(0, 0.3, (n, 3)) * np.array([0.01, 0.01, 1])
        off = np.array([0, n], np.int64)
        b0 = ref_lib.wheel_preintegrate(off, steps)[0]
        b1 = oracle.wheel_preintegrate(params, off, steps)[0]
        assert rel(b1[0:12], b0[0:12]) < 1e-13 and rel(b1[12:15], b0[12:15]) < 1e-12
    # standing still: the 0.005^2 floors of get_preintegraption_result (wheel_odom_preintegration.h:115-116)
    steps = np.zeros((3, 7))
    steps[:, 0] = 0.02
    b0, b1 = ref_lib.wheel_preintegrate([0, 3], steps)[0], oracle.wheel_preintegrate(params, [0, 3], steps)[0]
    assert rel(b1, b0) < 1e-13


def test_imu_factor(oracle, params):
    worst_r = worst_j = 0.0
    for k in range(30):
        samples, bias = make_imu_blob(oracle, params, 20)
        blob = ref_lib.imu_preintegrate([0, 20], samples, bias[None])[0]
        si = rand_state(corridor_like_pose() if k % 2 else rand_pose())
        sj = si.copy()
        sj[0:3] += si[6:9] * blob[465] + RNG.normal(0, 0.02, 3)
        sj[3:6] = (Rotation.from_rotvec(si[3:6]) * Rotation.from_rotvec(blob[6:9] + RNG.normal(0, 0.01, 3))).as_rotvec()
        sj[6:9] += RNG.normal(0, 0.1, 3)
        sj[9:15] += RNG.normal(0, 1e-3, 6)
        r0, J0 = ref_lib.eval_imu_factor(blob, si, sj)
        r1, J1 = oracle.eval_imu_factor(params, blob, si, sj)
        worst_r, worst_j = max(worst_r, rel(r1, r0)), max(worst_j, rel(J1, J0))
    assert worst_r < 1e-12 and worst_j < 1e-11, (worst_r, worst_j)


def test_wheel_factor_all_branches(oracle, params):
    """wheel_factor.h:45-70 has three pairs of branches (translation / direction / rotation below their thresholds).
    The degenerate branches take the norm of a tiny vector; the inputs keep that vector >= 1e-5 so that the comparison is
    about the formulas, not about the direction of rounding noise (at exactly zero the reference's own Jet Jacobian is NaN)."""
    T_io = params_T(params, "T_imu_to_wheel")
    Tio4 = np.eye(4)
    Tio4[:3, :4] = T_io
    branches = set()
    worst_r = worst_j = 0.0
    for k in range(80):
        mode = k % 4
        steps = np.zeros((5, 7))
        steps[:, 0] = 0.02
        steps[:, 1] = RNG.uniform(0.2, 1.0) if mode != 1 else RNG.uniform(2e-4, 5e-4)      # forward speed: |op| >= or < 1e-4
        steps[:, 6] = RNG.normal(0, 0.5) if mode != 2 else RNG.uniform(2e-4, 5e-3)         # yaw rate: |oq| >= or < 1e-3
        blob = ref_lib.wheel_preintegrate([0, 5], steps)[0]
        pi_ = corridor_like_pose()
        Ri = Rotation.from_rotvec(pi_[3:6]).as_matrix()
        dT = np.eye(4)
        dT[:3, :4] = blob[0:12].reshape(3, 4)
        if mode == 3:   # frame j barely moved: |p| < 1e-4 and |q| < 1e-3 whatever the wheels measured
            dT = np.eye(4)
            dT[:3, :3] = Rotation.from_rotvec([0, 0, RNG.uniform(2e-5, 8e-4)]).as_matrix()
            dT[:3, 3] = [RNG.uniform(2e-5, 8e-5), RNG.uniform(-2e-5, 2e-5), 0]
        Two_i = np.eye(4)
        Two_i[:3, :3], Two_i[:3, 3] = Ri @ T_io[:, :3], Ri @ T_io[:, 3] + pi_[0:3]
        Twi_j = Two_i @ dT @ np.linalg.inv(Tio4)
        noise = RNG.normal(0, 1e-3, 3) if mode == 0 else np.zeros(3)
        pj_ = np.concatenate([Twi_j[:3, 3] + noise, Rotation.from_matrix(Twi_j[:3, :3]).as_rotvec()])
        r0, J0 = ref_lib.eval_wheel_factor(blob, pi_, pj_)
        r1, J1 = oracle.eval_wheel_factor(params, blob, pi_, pj_)
        assert np.all(np.isfinite(J0))
        worst_r = max(worst_r, np.abs(r1 - r0).max() / max(np.abs(r0).max(), 1.0))   # whitened residuals: O(1) is one sigma
        worst_j = max(worst_j, max(np.abs(J1[i] - J0[i]).max() / np.abs(J0[i]).max() for i in range(3)))
        dTm = blob[0:12].reshape(3, 4)
        branches.add((np.hypot(*dTm[:2, 3]) < 1e-4, np.linalg.norm(Rotation.from_matrix(dTm[:, :3]).as_rotvec()) < 1e-3, mode == 3))
    assert len(branches) >= 4, branches
    assert worst_r < 1e-11 and worst_j < 1e-9, (worst_r, worst_j)


def test_ground_and_prior_factors(oracle, params):
    for k in range(40):
        pose = corridor_like_pose()
        pose[3:6] = (Rotation.from_rotvec(pose[3:6]) * Rotation.from_rotvec(RNG.normal(0, 1e-3 if k % 2 else 0.05, 3))).as_rotvec()
        r0, J0 = ref_lib.eval_ground_factors(pose)
        r1, J1 = oracle.eval_ground_factors(params, pose)
        assert rel(r1, r0) < 1e-10 and rel(J1, J0) < 1e-9, (k, r0, r1)
    for _ in range(10):
        X0, J, s = rand_state(), RNG.normal(size=(15, 15)), rand_state()
        r0, J0 = ref_lib.eval_prior_factor(X0, J, s)
        r1, J1 = oracle.eval_prior_factor(X0, J, s)
        assert rel(r1, r0) < 1e-13 and rel(J1, J0) < 1e-13
        assert rel(r0, J @ (s - X0)) < 1e-13            # marginalization_factor.h:50: no linearized_R term


def test_edge_factor_and_noise(oracle):
    """edge_factor.h:79-126 and edge_noise (:14-26), whose J(1, 2) = 1 / sigma_p(1) leaves J(1, 1) = 1."""
    from test_oracle_pose_graph import edge_noise_J
    for k in range(30):
        pi_, pj_ = rand_pose(), rand_pose()
        tf = np.eye(4)[:3]
        tf[:, :3], tf[:, 3] = Rotation.from_rotvec(rand_rotvec(1.0)).as_matrix(), RNG.normal(0, 1, 3)
        w = RNG.uniform(0.5, 10)
        r0, J0, Jn = ref_lib.eval_edge_factor(tf, w, pi_, pj_)
        assert np.array_equal(Jn, edge_noise_J())
        r1, J1 = oracle.eval_edge_factor(tf, w, Jn, pi_, pj_)
        assert rel(r1, r0) < 1e-11 and rel(J1, J0) < 1e-10


def test_golden_factors_against_reference_text():
    """tests/golden/factors.npz was written by the oracle in round 1; the reference text reproduces it."""
    z = np.load(os.path.join(GOLD, "factors.npz"))
    r, J = ref_lib.eval_imu_factor(z["imu_blob"], z["state_i"], z["state_j"])
    assert rel(z["imu_res"], r) < 1e-11 and rel(z["imu_jac"], J) < 1e-11
    r, J = ref_lib.eval_wheel_factor(z["wheel_blob"], z["state_i"][:6], z["state_j"][:6])
    assert rel(z["wheel_res"], r) < 1e-11 and rel(z["wheel_jac"], J) < 1e-10
    l = z["laser_lines"]
    r, J = ref_lib.eval_laser_factor(l[0], l[1], l[2], l[3], z["state_i"][:6], z["laser_pose_j"])
    assert rel(z["laser_res"], r) < 1e-12 and rel(z["laser_jac"], J) < 1e-11
