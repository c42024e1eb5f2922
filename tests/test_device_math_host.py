"""Host build of the product's device math header (csrc/lv_math.cuh) checked against the oracle.

The header is `__host__ __device__`; tests/native/lv_math_host.cpp compiles it with g++ so that the closed-form /
dual-number Jacobian columns the CUDA kernels use can be verified here without a GPU.  This is a formula check,
not a CPU fallback of the product.  Tolerances: values 1e-12, Jacobians 1e-10 (relative to the largest entry)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200.params import params_T
from test_oracle_factors import RNG, corridor_like_pose, make_imu_blob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Consts(C.Structure):
    _fields_ = [("T_il", C.c_double * 12), ("T_io", C.c_double * 12), ("g", C.c_double), ("laser_sqrt_info", C.c_double),
                ("ground_p_sqrt_info", C.c_double), ("ground_q_sqrt_info", C.c_double), ("Q", C.c_double * 12),
                ("wheel_cov", C.c_double * 3), ("huber_delta", C.c_double)]


@pytest.fixture(scope="module")
def lvm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lvm") / "liblvm_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(ROOT, "tests", "native", "lv_math_host.cpp"), "-o", out])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def consts(params):
    c = Consts()
    c.T_il[:] = list(params.T_imu_to_laser)
    c.T_io[:] = list(params.T_imu_to_wheel)
    c.g = params.g
    c.laser_sqrt_info = 1.0 / params.line_to_line_sigma
    c.ground_p_sqrt_info = 1.0 / params.manifold_p_sigma
    c.ground_q_sqrt_info = 1.0 / params.manifold_q_sigma
    return c


def d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_exp_log(lvm, oracle):
    for _ in range(100):
        v = RNG.normal(size=3)
        v = v / np.linalg.norm(v) * RNG.uniform(0, 3.1)
        R = np.zeros(9)
        lvm.lvm_exp(d(v), d(R))
        np.testing.assert_allclose(R.reshape(3, 3), oracle.exp_so3(v), rtol=0, atol=1e-15)
        w = np.zeros(3)
        lvm.lvm_log(d(R), d(w))
        np.testing.assert_allclose(w, oracle.log_SO3(R.reshape(3, 3)), rtol=0, atol=1e-14)
    out = np.zeros(3)
    lvm.lvm_so3_plus(d(np.array([3.0, 0.1, 0.0])), d(np.array([0.4, 0.0, 0.0])), d(out))
    np.testing.assert_allclose(out, oracle.so3_plus([3.0, 0.1, 0.0], [0.4, 0.0, 0.0]), rtol=1e-15)


def test_imu_columns_match_oracle(lvm, oracle, params, consts):
    for _ in range(10):
        blob, _, _ = make_imu_blob(oracle, params)
        si = np.concatenate([corridor_like_pose(), RNG.normal(0, 0.5, 3), RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])
        sj = si + np.concatenate([RNG.normal(0, 0.05, 3), RNG.normal(0, 0.03, 3), RNG.normal(0, 0.1, 3), RNG.normal(0, 1e-3, 6)])
        r_raw, J_raw = np.zeros(15), np.zeros((15, 30))
        lvm.lvm_imu(C.byref(consts), d(blob), d(si), d(sj), d(r_raw), d(J_raw))
        S = blob[240:465].reshape(15, 15)
        res, jac = oracle.eval_imu_factor(params, blob, si, sj)
        np.testing.assert_allclose(S @ r_raw, res, rtol=1e-11, atol=1e-12 * np.abs(res).max())
        np.testing.assert_allclose(S @ J_raw, jac, rtol=1e-9, atol=1e-11 * np.abs(jac).max())


def test_wheel_and_ground_match_oracle(lvm, oracle, params, consts):
    T_io = params_T(params, "T_imu_to_wheel")
    for case in range(10):
        steps = np.zeros((2, 7)); steps[:, 0] = 0.05
        if case < 7:
            steps[:, 1:4] = [0.7, 0.01, 0.0] + RNG.normal(0, 0.02, (2, 3)); steps[:, 4:7] = [0, 0, 0.3] + RNG.normal(0, 0.02, (2, 3))
        else:
            steps[:, 1:4] = RNG.normal(0, 1e-5, (2, 3)); steps[:, 4:7] = RNG.normal(0, 1e-4, (2, 3))
        blob = oracle.wheel_preintegrate(params, [0, 2], steps)[0]
        pi_ = corridor_like_pose()
        pj_ = pi_ + np.concatenate([RNG.normal(0, 0.04, 3), RNG.normal(0, 0.01, 3)])
        res, J = np.zeros(3), np.zeros((3, 12))
        lvm.lvm_wheel(C.byref(consts), d(blob), d(pi_), d(pj_), d(res), d(J))
        ores, ojac = oracle.eval_wheel_factor(params, blob, pi_, pj_)
        np.testing.assert_allclose(res, ores, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(J, ojac, rtol=1e-9, atol=1e-10 * np.abs(ojac).max())
        gres, gJ = np.zeros(2), np.zeros((2, 6))
        lvm.lvm_ground(C.byref(consts), d(pi_), d(gres), d(gJ))
        ogres, ogjac = oracle.eval_ground_factors(params, pi_)
        np.testing.assert_allclose(gres, ogres, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(gJ, ogjac, rtol=1e-9, atol=1e-10 * np.abs(ogjac).max())


def test_laser_frame_table_reproduces_laser_jacobian(lvm, oracle, params, consts):
    """d = n.(M c + t) - n.A2 and d(d)/d(theta_k) = n.(B_k c + b_k): the per-point algebra of the scan-match kernel."""
    for _ in range(10):
        pose_i, pose_j = corridor_like_pose(), corridor_like_pose()
        pose_j[0:3] = pose_i[0:3] + RNG.normal(0, 0.3, 3)
        tab_i, tab_j = np.zeros(24), np.zeros(24)
        lvm.lvm_frame_table(C.byref(consts), d(pose_i), d(tab_i))
        lvm.lvm_frame_table(C.byref(consts), d(pose_j), d(tab_j))
        a1, a2, c = RNG.uniform(-5, 5, 2), RNG.uniform(-5, 5, 2), RNG.uniform(-5, 5, 2)
        Mi, ti = tab_i[0:4].reshape(2, 2), tab_i[4:6]
        Mj, tj = tab_j[0:4].reshape(2, 2), tab_j[4:6]
        A1, A2 = Mi @ a1 + ti, Mi @ a2 + ti
        u = (A2 - A1) / np.linalg.norm(A2 - A1)
        n = np.array([-u[1], u[0]])
        dist = n @ (Mj @ c + tj - A2)
        w = 1.7
        res, jac = oracle.eval_laser_point(params, a1, a2, c, w, pose_i, pose_j)
        k = w / params.line_to_line_sigma
        assert abs(k * abs(dist) - res[0]) < 1e-9 * max(1.0, abs(res[0]))
        s = np.sign(dist)
        jth = np.array([n @ (tab_j[6 + 6 * q:10 + 6 * q].reshape(2, 2) @ c + tab_j[10 + 6 * q:12 + 6 * q]) for q in range(3)])
        want = k * s * np.concatenate([n, [0.0], jth])
        np.testing.assert_allclose(want, jac[0, 6:12], rtol=1e-9, atol=1e-8 * np.abs(jac).max())
        # reference-side pose (free in the init topology): d(d)/d(theta_i,k) = -(u.(C-A2)) (n.dDelta_k)/L - n.dA2_k
        L_ = np.linalg.norm(A2 - A1)
        tt = u @ (Mj @ c + tj - A2)
        jti = []
        for q in range(3):
            Bk, bk = tab_i[6 + 6 * q:10 + 6 * q].reshape(2, 2), tab_i[10 + 6 * q:12 + 6 * q]
            jti.append(-tt * (n @ (Bk @ (a2 - a1))) / L_ - n @ (Bk @ a2 + bk))
        want_i = k * s * np.concatenate([-n, [0.0], jti])
        np.testing.assert_allclose(want_i, jac[0, 0:6], rtol=1e-9, atol=1e-8 * np.abs(jac).max())


@pytest.mark.parametrize("entry", ["lvm_item", "lvm_item_analytic"])
def test_unified_item_matches_oracle(lvm, oracle, params, consts, entry):
    """One (frame i-1, frame i) item: IMU + wheel + ground residuals and Jacobians, through the dual-number path
    (item_residuals<Dual>) and through the closed-form path factor_kernel uses (item_phase1/2 + item_column)."""
    for trial in range(12):
        blob, _, _ = make_imu_blob(oracle, params)
        steps = np.zeros((2, 7)); steps[:, 0] = 0.05
        if trial < 9:
            steps[:, 1:4] = [0.7, 0.01, 0.0] + RNG.normal(0, 0.02, (2, 3)); steps[:, 4:7] = [0, 0, 0.3] + RNG.normal(0, 0.02, (2, 3))
        else:   # standing still: the degenerate branches of wheel_factor.h:45-70
            steps[:, 1:4] = RNG.normal(0, 1e-5, (2, 3)); steps[:, 4:7] = RNG.normal(0, 1e-4, (2, 3))
        wblob = oracle.wheel_preintegrate(params, [0, 2], steps)[0]
        si = np.concatenate([corridor_like_pose(), RNG.normal(0, 0.5, 3), RNG.normal(0, 0.02, 3), RNG.normal(0, 0.002, 3)])
        dpose = np.concatenate([RNG.normal(0, 0.05, 3), RNG.normal(0, 0.03, 3)]) if trial < 9 else np.concatenate([RNG.normal(0, 2e-5, 3), RNG.normal(0, 2e-4, 3)])
        sj = si + np.concatenate([dpose, RNG.normal(0, 0.1, 3), RNG.normal(0, 1e-3, 6)])
        r_imu, J_imu = np.zeros(15), np.zeros((15, 30))
        r_w, J_w, r_g, J_g = np.zeros(3), np.zeros((3, 30)), np.zeros(2), np.zeros((2, 30))
        getattr(lvm, entry)(C.byref(consts), d(blob), d(wblob), d(si), d(sj), d(r_imu), d(J_imu), d(r_w), d(J_w), d(r_g), d(J_g))
        S = blob[240:465].reshape(15, 15)
        res, jac = oracle.eval_imu_factor(params, blob, si, sj)
        np.testing.assert_allclose(S @ r_imu, res, rtol=1e-11, atol=1e-12 * np.abs(res).max())
        np.testing.assert_allclose(S @ J_imu, jac, rtol=1e-9, atol=1e-11 * np.abs(jac).max())
        wres, wjac = oracle.eval_wheel_factor(params, wblob, si[:6], sj[:6])
        np.testing.assert_allclose(r_w, wres, rtol=1e-11, atol=1e-9)
        pose_cols = [0, 1, 2, 3, 4, 5, 15, 16, 17, 18, 19, 20]
        np.testing.assert_allclose(J_w[:, pose_cols], wjac, rtol=1e-9, atol=1e-10 * np.abs(wjac).max())
        assert np.all(np.delete(J_w, pose_cols, axis=1) == 0)
        gres, gjac = oracle.eval_ground_factors(params, sj[:6])
        np.testing.assert_allclose(r_g, gres, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(J_g[:, 15:21], gjac, rtol=1e-9, atol=1e-10 * np.abs(gjac).max())
