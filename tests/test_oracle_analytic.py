"""The "analytic" CPU-baseline flavour of the oracle (closed-form Jacobian of the scan-point factor)
against the faithful Jet flavour: same residual, same Jacobian, same solve."""
import time

import numpy as np

import lvio2d_b200 as L


def test_closed_form_laser_jacobian_equals_the_jet_jacobian(oracle):
    P = L.corridor_params()
    g = np.random.default_rng(5)
    for _ in range(200):
        a1, a2 = g.uniform(-8, 8, 2), g.uniform(-8, 8, 2)
        c = g.uniform(-8, 8, 2)
        pose_i = np.r_[g.uniform(-3, 3, 3), g.normal(0, 0.6, 3)]
        pose_j = np.r_[g.uniform(-3, 3, 3), g.normal(0, 0.6, 3)]
        w = g.uniform(0.5, 3.0)
        r0, J0 = oracle.eval_laser_point(P, a1, a2, c, w, pose_i, pose_j)
        r1, J1 = oracle.eval_laser_point_analytic(P, a1, a2, c, w, pose_i, pose_j)
        assert abs(r0[0] - r1[0]) <= 1e-12 * max(1.0, abs(r0[0]))
        assert np.abs(J0 - J1).max() <= 1e-9 * max(1.0, np.abs(J0).max())


def test_analytic_flavour_solves_like_the_jet_flavour_and_is_faster(oracle):
    for sb, iters in ((L.synth.make_batch(2, 42, n_frames=5, beams=300, fov_deg=270.0), 10), (L.synth.config_init(1, n_frames=5), 10)):
        P = L.corridor_params(max_iters=iters)
        hb = oracle.preintegrate_batch(P, sb)
        t0 = time.perf_counter()
        want, s0 = oracle.solve(P, hb)
        t1 = time.perf_counter()
        got, s1 = oracle.solve(P, hb, analytic=True)
        t2 = time.perf_counter()
        assert np.array_equal(s0["iterations"], s1["iterations"])
        assert np.abs(got - want).max() < 1e-8
        assert (t2 - t1) < (t1 - t0)
