"""tests/golden/ref_text.npz holds outputs of the REFERENCE'S OWN source text (written by scripts/make_golden_ref.py
through oracle/_ref/libref.so in the container that has /root/reference).  These tests need neither the reference tree
nor libref.so: they pin the oracle (CPU) and the CUDA path (GPU, through the C ABI) to that fixture — residuals and
Jacobians of every factor, both preintegrators, solver::solve in fast_mode on a 2- and a 6-frame window,
solver::marginalization, laser_manager::spawn_scan and do_match."""
import copy
import os

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200 import replay
from lvio2d_b200.solver import Solver

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_text.npz"))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


class _OracleHooks:
    """the oracle behind the method names of solver.Context (the factor hooks take `params` first there)"""

    def __init__(self, oracle, P):
        self.o, self.P = oracle, P

    def eval_laser_factor(self, *a): return self.o.eval_laser_factor(self.P, *a)
    def eval_imu_factor(self, *a): return self.o.eval_imu_factor(self.P, *a)
    def eval_wheel_factor(self, *a): return self.o.eval_wheel_factor(self.P, *a)
    def eval_ground_factors(self, *a): return self.o.eval_ground_factors(self.P, *a)
    def imu_preintegrate(self, *a): return self.o.imu_preintegrate(self.P, *a)
    def wheel_preintegrate(self, *a): return self.o.wheel_preintegrate(self.P, *a)


def check_factors(be, tol_r, tol_j):
    n = len(G["laser_res"])
    for k in range(n):
        r, J = be.eval_laser_factor(*G["laser_ends"][k], G["laser_pose_i"][k], G["laser_pose_j"][k])
        assert rel(r, G["laser_res"][k]) < tol_r and rel(J, G["laser_jac"][k]) < tol_j, ("laser", k)
        r, J = be.eval_imu_factor(G["imu_blob"][k], G["imu_si"][k], G["imu_sj"][k])
        assert rel(r, G["imu_res"][k]) < tol_r and rel(J, G["imu_jac"][k]) < tol_j, ("imu", k)
        r, J = be.eval_wheel_factor(G["wheel_blob"][k], G["wheel_pose_i"][k], G["wheel_pose_j"][k])
        assert np.abs(np.asarray(r) - G["wheel_res"][k]).max() / max(np.abs(G["wheel_res"][k]).max(), 1.0) < tol_r, ("wheel", k)
        assert rel(J, G["wheel_jac"][k]) < tol_j, ("wheel", k)
        r, J = be.eval_ground_factors(G["ground_pose"][k])
        assert rel(r, G["ground_res"][k]) < 1e-9 and rel(J, G["ground_jac"][k]) < 1e-8, ("ground", k)
        blob = be.imu_preintegrate(np.array([0, 20], np.int64), G["imu_samples"][k], G["imu_bias"][k][None])[0]
        assert rel(blob[0:240], G["imu_blob"][k][0:240]) < 1e-11 and rel(blob[240:465], G["imu_blob"][k][240:465]) < 1e-7
        wb = be.wheel_preintegrate(np.array([0, 5], np.int64), G["wheel_steps"][k])[0]
        assert rel(wb, G["wheel_blob"][k]) < 1e-12


def check_solver(P, make_solver, oracle):
    sb = replay.make_sequence(5, n_frames=8, params=P)
    hb = oracle.preintegrate_batch(P, sb)
    frames = replay.frames_of(sb, hb["imu"], hb["wheel"])
    for n in (2, 6):
        s = make_solver(True)
        a = copy.deepcopy(frames[:n])
        s.solve(a)
        got = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in a])
        want, summ = G[f"solve{n}_states"], G[f"solve{n}_summary"]
        assert int(s.last_summary["iterations"][0]) == int(summ[0]) and int(s.last_summary["termination"][0]) == int(summ[1]), n
        assert float(s.last_summary["initial_cost"][0]) == pytest.approx(summ[2], rel=1e-10)
        assert float(s.last_summary["final_cost"][0]) == pytest.approx(summ[3], rel=1e-7)
        assert np.abs(got - want).max() < 1e-8, (n, np.abs(got - want).max())
    s = make_solver(False)
    a = copy.deepcopy(frames[:3])
    s.marginalization(a)
    assert np.abs(s.linearized_X - G["marg_X0"]).max() < 1e-15
    JTJ = s.linearized_jacobians.T @ s.linearized_jacobians
    assert np.abs(JTJ - G["marg_JTJ"]).max() / np.abs(G["marg_JTJ"]).max() < 1e-7


def check_front_end(be, P):
    lp = L.corridor_line_params()
    off, pts = L.synth.make_scan_batch(3, 11, beams=721, range_sigma=0.004)
    n, lines, abc, rng = be.extract_lines(lp, off, pts, max_lines=96)
    for k in range(3):
        want = G[f"scan{k}_lines"]
        assert n[k] == len(want), k
        assert np.abs(lines[k, :n[k]] - want[:, [0, 1, 3, 4]]).max() < 1e-9, k
    q3, l2 = G["match_points2"], G["match_lines2"]
    n2, lines2, _, _ = be.extract_lines(lp, np.array([0, len(q3)], np.int64), q3[:, :2], max_lines=96)
    assert n2[0] == len(l2) and np.abs(lines2[0, :n2[0]] - l2[:, [0, 1, 3, 4]]).max() < 1e-9
    for kk in (0, 1):
        nm, m = be.match_lines(lp, n[0:1], lines[0:1], n2, lines2, G["match_pose1"][None], G["match_pose2"][None], kk=kk,
                               point_offset1=np.array([0, off[1]], np.int64), points1=pts[off[0]:off[1]], index_range1=rng[0:1])
        want = G[f"match_pairs_kk{kk}"]
        assert nm[0] == len(want) and np.array_equal(m[0, :nm[0]], want), kk


def test_oracle_reproduces_the_reference_text_fixture(oracle):
    P = L.corridor_params()
    check_factors(_OracleHooks(oracle, P), 1e-11, 1e-9)
    check_solver(L.corridor_params(fast_mode=True), lambda fast: Solver(L.corridor_params(fast_mode=fast), fast_mode=fast,
                                                                        ctx=oracle.OracleContext(L.corridor_params(fast_mode=fast))), oracle)

    class FE:
        extract_lines = staticmethod(oracle.extract_lines)

        @staticmethod
        def match_lines(lp, *a, **kw):
            return oracle.match_lines(P, lp, *a, **kw)

    check_front_end(FE, P)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_text_fixture(oracle):
    from lvio2d_b200.solver import Context

    P = L.corridor_params()
    with Context(P) as c:
        check_factors(c, 1e-9, 1e-8)
        check_front_end(c, P)
    solvers = []

    def make(fast):
        s = Solver(L.corridor_params(fast_mode=fast), fast_mode=fast)
        solvers.append(s)
        return s

    try:
        check_solver(L.corridor_params(fast_mode=True), make, oracle)
    finally:
        for s in solvers:
            s.close()
