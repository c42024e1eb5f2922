"""Pins the CPU oracle (oracle/) — the reference ships no tests, so the pins are an independent torch-float64
re-derivation (tests/independent_ref.py), scipy's Rotation and central finite differences.
Tolerances: values 1e-12 relative, Jacobians 1e-9 relative."""
import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

import independent_ref as ref
import lvio2d_b200 as L
from lvio2d_b200.params import params_T

RNG = np.random.default_rng(7)


def rand_pose(scale_q=2.0):
    q = RNG.normal(size=3)
    q = q / np.linalg.norm(q) * RNG.uniform(0.1, scale_q)
    return np.concatenate([RNG.uniform(-5, 5, 3), q])


def corridor_like_pose():
    """|q| ~ 2 rad like the corridor extrinsics give."""
    T_io = params_T(L.corridor_params(), "T_imu_to_wheel")
    Rwb = Rotation.from_euler("zyx", [RNG.uniform(-3, 3), RNG.normal(0, 2e-3), RNG.normal(0, 2e-3)]).as_matrix()
    Rwi = Rwb @ T_io[:, :3].T
    return np.concatenate([RNG.uniform(-5, 5, 3), Rotation.from_matrix(Rwi).as_rotvec()])


def close(a, b, rtol, atol=1e-12):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_exp_log_against_scipy(oracle):
    for _ in range(200):
        v = RNG.normal(size=3)
        v = v / np.linalg.norm(v) * RNG.uniform(0, np.pi * 0.999)
        R = oracle.exp_so3(v)
        close(R, Rotation.from_rotvec(v).as_matrix(), 1e-12, 1e-14)
        close(oracle.log_SO3(R), v, 1e-9, 1e-12)
    close(oracle.exp_so3(np.zeros(3)), np.eye(3), 0, 0)
    # the corridor initial orientation has |q| = 2.078 rad
    T_io = params_T(L.corridor_params(), "T_imu_to_wheel")
    q0 = oracle.log_SO3(T_io[:, :3].T)
    assert abs(np.linalg.norm(q0) - 2.078) < 2e-3


def test_normalize_so3_wraps(oracle):
    v = np.array([0.3, -0.2, 0.9])
    for ang in (0.5, 3.0, np.pi + 0.1, 2 * np.pi + 0.4, 5 * np.pi - 0.2):
        a = v / np.linalg.norm(v) * ang
        out = oracle.normalize_so3(a)
        assert np.linalg.norm(out) <= np.pi + 1e-12
        close(Rotation.from_rotvec(out).as_matrix(), Rotation.from_rotvec(a).as_matrix(), 1e-9, 1e-12)
        if ang <= np.pi:
            assert np.array_equal(out, a)
    close(oracle.so3_plus([3.0, 0, 0], [0.3, 0, 0]), [3.3 - 2 * np.pi, 0, 0], 1e-12)


def test_dis_from_line(oracle):
    assert abs(oracle.dis_from_line([0.5, 2.0, 0], [0, 0, 0], [1, 0, 0]) - 2.0) < 1e-15
    assert abs(oracle.dis_from_line([3.0, -1.5, 0], [0, 0, 0], [1, 0, 0]) - 1.5) < 1e-15  # infinite line
    for _ in range(50):
        p, p1, p2 = RNG.normal(size=3), RNG.normal(size=3), RNG.normal(size=3)
        u = (p2 - p1) / np.linalg.norm(p2 - p1)
        want = np.linalg.norm(np.cross(p - p1, u))
        assert abs(oracle.dis_from_line(p, p1, p2) - want) < 1e-12


def test_laser_factor_matches_independent(oracle, params):
    T_il = params_T(params, "T_imu_to_laser")
    for _ in range(25):
        pi_, pj_ = corridor_like_pose(), corridor_like_pose()
        pj_[0:3] = pi_[0:3] + RNG.normal(0, 0.3, 3)
        l1_p1, l1_p2 = np.append(RNG.uniform(-6, 6, 2), 0), np.append(RNG.uniform(-6, 6, 2), 0)
        l2_p1, l2_p2 = np.append(RNG.uniform(-6, 6, 2), 0), np.append(RNG.uniform(-6, 6, 2), 0)
        res, jac = oracle.eval_laser_factor(params, l1_p1, l1_p2, l2_p1, l2_p2, pi_, pj_)
        t = [torch.tensor(x) for x in (l1_p1, l1_p2, l2_p1, l2_p2)]
        fn = lambda a, b: ref.laser_pair_residual(a, b, *t, T_il, params.line_to_line_sigma)
        want = fn(torch.tensor(pi_), torch.tensor(pj_)).numpy()
        close(res, want, 1e-11, 1e-9)
        close(jac, ref.jac(fn, pi_, pj_), 1e-9, 1e-7)
        # point form == one residual of the pair form
        w = np.sqrt(min(np.linalg.norm(l1_p1 - l1_p2), np.linalg.norm(l2_p1 - l2_p2)) / 2 / 0.02)
        r1, j1 = oracle.eval_laser_point(params, l1_p1[:2], l1_p2[:2], l2_p1[:2], w, pi_, pj_)
        close(r1[0], res[0], 1e-14)
        close(j1[0], jac[0], 1e-14, 1e-9)
        # the z column of d/dp is exactly zero (P = diag(1,1,0))
        assert jac[0, 2] == 0.0 and jac[0, 8] == 0.0


def make_imu_blob(oracle, params, n_samples=20, dt=0.005):
    samples = np.zeros((n_samples, 7))
    samples[:, 0] = dt
    samples[:, 1:4] = np.array([0.1, -9.7, 0.3]) + RNG.normal(0, 0.2, (n_samples, 3))
    samples[:, 4:7] = RNG.normal(0, 0.3, (n_samples, 3))
    bias0 = np.concatenate([RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])[None]
    return oracle.imu_preintegrate(params, [0, n_samples], samples, bias0)[0], samples, bias0


def test_imu_factor_matches_independent(oracle, params):
    for _ in range(10):
        blob, _, _ = make_imu_blob(oracle, params)
        si = np.concatenate([corridor_like_pose(), RNG.normal(0, 0.5, 3), RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])
        sj = si + np.concatenate([RNG.normal(0, 0.05, 3), RNG.normal(0, 0.03, 3), RNG.normal(0, 0.1, 3), RNG.normal(0, 1e-3, 6)])
        res, jac = oracle.eval_imu_factor(params, blob, si, sj)
        fn = lambda a, b: ref.imu_residual(a, b, blob, params.g)
        want = fn(torch.tensor(si), torch.tensor(sj)).numpy()
        close(res, want, 1e-9, 1e-6 * np.abs(want).max())
        wj = ref.jac(fn, si, sj)
        close(jac, wj, 1e-8, 1e-9 * np.abs(wj).max())


def test_wheel_factor_matches_independent(oracle, params):
    T_io = params_T(params, "T_imu_to_wheel")
    for case in range(12):
        steps = np.zeros((2, 7))
        steps[:, 0] = 0.05
        if case < 8:
            steps[:, 1:4] = [0.7, 0.01, 0.0] + RNG.normal(0, 0.02, (2, 3))
            steps[:, 4:7] = [0.0, 0.0, 0.3] + RNG.normal(0, 0.02, (2, 3))
        elif case < 10:  # standing still: the len < 1e-4 / |q| < 1e-3 branches
            steps[:, 1:4] = RNG.normal(0, 1e-5, (2, 3))
            steps[:, 4:7] = RNG.normal(0, 1e-4, (2, 3))
        else:            # pure rotation
            steps[:, 1:4] = RNG.normal(0, 1e-5, (2, 3))
            steps[:, 4:7] = [0.0, 0.0, 0.4]
        blob = oracle.wheel_preintegrate(params, [0, 2], steps)[0]
        pi_ = corridor_like_pose()
        Ri = Rotation.from_rotvec(pi_[3:6]).as_matrix()
        # pose_j: pose_i moved by roughly the odometry increment (+ noise), expressed through the extrinsic
        dR, dt = blob[:12].reshape(3, 4)[:, :3], blob[:12].reshape(3, 4)[:, 3]
        Roi = Ri @ T_io[:, :3]
        toi = Ri @ T_io[:, 3] + pi_[0:3]
        Roj = Roi @ dR @ Rotation.from_rotvec(RNG.normal(0, 2e-3, 3)).as_matrix()
        toj = toi + Roi @ (dt + RNG.normal(0, 2e-3, 3))
        Rj = Roj @ T_io[:, :3].T
        pj_ = np.concatenate([toj - Rj @ T_io[:, 3], Rotation.from_matrix(Rj).as_rotvec()])
        res, jac = oracle.eval_wheel_factor(params, blob, pi_, pj_)
        fn = lambda a, b: ref.wheel_residual(a, b, blob, T_io)
        want = fn(torch.tensor(pi_), torch.tensor(pj_)).numpy()
        close(res, want, 1e-8, 1e-7 * max(1.0, np.abs(want).max()))
        wj = ref.jac(fn, pi_, pj_)
        close(jac, wj, 1e-6, 1e-8 * np.abs(wj).max())


def test_ground_factors_match_independent_and_fd(oracle, params):
    T_io = params_T(params, "T_imu_to_wheel")
    for _ in range(20):
        pose = corridor_like_pose()
        pose[2] = 0.92 + RNG.normal(0, 0.01)
        res, jac = oracle.eval_ground_factors(params, pose)
        fn = lambda a: ref.ground_residuals(a, T_io, params.manifold_p_sigma, params.manifold_q_sigma)
        want = fn(torch.tensor(pose)).numpy()
        close(res, want, 1e-8, 1e-7)
        close(jac, ref.jac(fn, pose), 1e-7, 1e-6)
        # central differences
        fd = np.zeros((2, 6))
        for k in range(6):
            h = 1e-6
            e = np.zeros(6); e[k] = h
            fd[:, k] = (oracle.eval_ground_factors(params, pose + e)[0] - oracle.eval_ground_factors(params, pose - e)[0]) / (2 * h)
        close(jac, fd, 1e-5, 1e-3)


def test_prior_factor_is_linear(oracle):
    X0, J, x = RNG.normal(size=15), RNG.normal(size=(15, 15)), RNG.normal(size=15)
    res, jac = oracle.eval_prior_factor(X0, J, x)
    close(res, J @ (x - X0), 1e-13)   # linearized_R is not part of the residual (marginalization_factor.h:50)
    close(jac, J, 1e-15, 0)


def independent_imu_preintegration(params, samples, bias0):
    """numpy restatement of imu_preintegraption.h:170-208 through scipy rotations."""
    X = np.zeros(15); X[9:15] = bias0
    Jm = np.eye(15); Pm = np.eye(15) * 1e-5; Dt = 0.0
    Q = np.diag(np.concatenate([np.array(list(params.imu_noise_acc_sigma)) ** 2, np.array(list(params.imu_noise_gyro_sigma)) ** 2,
                                np.array(list(params.imu_bias_acc_sigma)) ** 2, np.array(list(params.imu_bias_gyro_sigma)) ** 2]))
    sk = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    for row in samples:
        dt, acc, gyro = row[0], row[1:4], row[4:7]
        a, b, gmm, ba, bw = X[0:3], X[3:6], X[6:9], X[9:12], X[12:15]
        Rz = Rotation.from_rotvec(gmm).as_matrix()
        Xn = X.copy()
        Xn[0:3] = a + b * dt + 0.5 * Rz @ (acc - ba) * dt * dt
        Xn[3:6] = b + Rz @ (acc - ba) * dt
        Xn[6:9] = (Rotation.from_rotvec(gmm) * Rotation.from_rotvec((gyro - bw) * dt)).as_rotvec()
        F = np.zeros((15, 15))
        F[0:3, 3:6] = np.eye(3)
        F[3:6, 6:9] = -Rz @ sk(acc - ba)
        F[3:6, 9:12] = -Rz
        F[6:9, 6:9] = -sk(gyro - ba)   # sic: last_ba (imu_preintegraption.h:192)
        F[6:9, 12:15] = -np.eye(3)
        G = np.zeros((15, 12))
        G[3:6, 0:3] = -Rz; G[6:9, 3:6] = -np.eye(3); G[9:12, 6:9] = np.eye(3); G[12:15, 9:12] = np.eye(3)
        F = np.eye(15) + F * dt
        Jm = F @ Jm
        Pm = F @ Pm @ F.T + (G * dt) @ Q @ (G * dt).T
        X = Xn; Dt += dt
    return X, Jm, Pm, Dt


def test_imu_preintegration(oracle, params):
    blob, samples, bias0 = make_imu_blob(oracle, params, n_samples=37)
    X, Jm, Pm, Dt = independent_imu_preintegration(params, samples, bias0[0])
    close(blob[:15], X, 1e-10, 1e-13)
    close(blob[15:240].reshape(15, 15), Jm, 1e-10, 1e-14)
    assert abs(blob[465] - Dt) < 1e-15
    S = blob[240:465].reshape(15, 15)
    assert np.allclose(S, np.triu(S)), "sqrt_inverse_P = L^T is upper triangular"
    close(S.T @ S @ Pm, np.eye(15), 0, 1e-6)
    want = np.linalg.cholesky(np.linalg.inv(Pm)).T
    close(S, want, 1e-6, 1e-6 * np.abs(want).max())


def test_wheel_preintegration(oracle, params):
    steps = np.zeros((3, 7)); steps[:, 0] = [0.05, 0.05, 0.02]
    steps[:, 1:4] = [0.8, 0.02, 0.0]; steps[:, 4:7] = [0.0, 0.0, 0.25]
    bad = np.array([[12.0, 1, 1, 1, 1, 1, 1], [-0.1, 1, 1, 1, 1, 1, 1]])  # dt >= 10 or <= 0: ignored (:143-147)
    blob = oracle.wheel_preintegrate(params, [0, 5], np.concatenate([steps, bad]))[0]
    T = np.eye(4)
    for row in steps:
        D = np.eye(4); D[:3, :3] = Rotation.from_rotvec(row[4:7] * row[0]).as_matrix(); D[:3, 3] = row[1:4] * row[0]
        T = T @ D
    close(blob[:12].reshape(3, 4), T[:3], 1e-12, 1e-15)
    dp, dq = T[:3, 3], Rotation.from_matrix(T[:3, :3]).as_rotvec()
    k = [max(dp @ dp, 2.5e-5), max(dp @ dp, 2.5e-5), max(dq @ dq, 2.5e-5)]
    want = [1 / np.sqrt(np.array(list(params.wheel_sigma))[i] ** 2 * k[i]) for i in range(3)]
    close(blob[12:15], want, 1e-12)
