"""Pins of the oracle's window-level code (linearisation assembly, LM loop, marginalisation) that do not need a GPU:
finite differences of the cost, scipy.optimize.least_squares, dense numpy Schur complements."""
import numpy as np
import pytest
from scipy.optimize import least_squares

import lvio2d_b200 as L
from lvio2d_b200 import abi


def small_window(oracle, P, seed=3, n_frames=3, beams=40, **kw):
    sb = L.synth.make_batch(1, seed, n_frames=n_frames, beams=beams, fov_deg=270.0, **kw)
    return sb, oracle.preintegrate_batch(P, sb)


def test_gradient_matches_finite_differences_of_cost(oracle):
    P = L.corridor_params(max_iters=10)
    for topo in ("tracking", "init"):
        sb, hb = small_window(oracle, P, topology=topo)
        x0 = hb["states"].copy()
        H, g, c = oracle.linearize(P, hb)
        assert abs(c[0] - oracle.cost(P, hb)[0]) < 1e-9 * c[0]
        fd = np.zeros_like(g[0])
        for k in range(g.shape[1]):
            f, col = divmod(k, 15)
            blk = 0 if col < 3 else (1 if col < 6 else (2 if col < 9 else 3))
            if (hb["const_mask"][f] >> blk) & 1:
                assert g[0, k] == 0.0 and np.all(H[0, k] == 0.0)
                continue
            h = 1e-7
            xp, xm = x0.copy(), x0.copy()
            xp.reshape(-1)[k] += h
            xm.reshape(-1)[k] -= h
            fd[k] = (oracle.cost(P, hb, xp)[0] - oracle.cost(P, hb, xm)[0]) / (2 * h)
        np.testing.assert_allclose(g[0], fd, rtol=2e-5, atol=2e-5 * np.abs(fd).max())
        assert np.allclose(H[0], H[0].T, rtol=0, atol=1e-9 * np.abs(H).max())
        assert np.linalg.eigvalsh(H[0]).min() > -1e-8 * np.abs(H).max()


def residual_vector(oracle, P, hb, x):
    """All residuals of the solver program of one small window, through the per-factor oracle entry points."""
    n = hb.n_frames
    x = x.reshape(n, 15)
    r = []
    po, lo = hb["point_offset"], hb["line_offset"]
    pts, lines = hb["points"].reshape(-1, 2), hb["lines"].reshape(-1, 4)
    for f in range(n):
        if hb["ref_frame"][f] < 0 and (hb["const_mask"][f] & 3) == 3:
            continue  # all parameter blocks constant: Ceres drops the residual block
        ref = hb["ref_pose"].reshape(-1, 6)[f] if hb["ref_frame"][f] < 0 else x[hb["ref_frame"][f], :6]
        for p in range(po[f], po[f + 1]):
            ln = lines[lo[f] + hb["point_line"][p]]
            w = 1.0 if hb["point_weight"] is None else hb["point_weight"][p]
            r.append(oracle.eval_laser_point(P, ln[0:2], ln[2:4], pts[p], w, ref, x[f, :6])[0][0])
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)
    for i in range(1, n):
        r += list(oracle.eval_imu_factor(P, imu[i - 1], x[i - 1], x[i])[0])
        r += list(oracle.eval_wheel_factor(P, wheel[i - 1], x[i - 1, :6], x[i, :6])[0])
    for f in range(n):
        if (hb["const_mask"][f] & 3) == 3:
            continue
        r += list(np.sqrt(hb.ground_multiplicity) * oracle.eval_ground_factors(P, x[f, :6])[0])
    if hb.prior_frame >= 0:
        r += list(oracle.eval_prior_factor(hb["prior_X0"].reshape(-1, 15)[0], hb["prior_J"].reshape(-1, 15, 15)[0], x[hb.prior_frame])[0])
    return np.array(r)


def test_lm_reaches_the_scipy_minimiser(oracle):
    """Gauge-fixed tracking window: the oracle's Ceres-style LM and scipy's trust-region solver must find the same
    minimiser (SURVEY.md §8c pin (ii))."""
    P = L.corridor_params(max_iters=200)
    P.function_tolerance, P.parameter_tolerance, P.gradient_tolerance = 1e-15, 1e-14, 1e-14
    sb, hb = small_window(oracle, P, seed=5, n_frames=3, beams=60)
    states, summ = oracle.solve(P, hb)
    c_or = oracle.cost(P, hb, states)[0]
    assert abs(0.5 * np.sum(residual_vector(oracle, P, hb, states) ** 2) - c_or) < 1e-8 * max(c_or, 1.0)
    free = np.ones(hb.n_frames * 15, bool)
    for f in range(hb.n_frames):
        for blk, (a, b) in enumerate([(0, 3), (3, 6), (6, 9), (9, 15)]):
            if (hb["const_mask"][f] >> blk) & 1:
                free[15 * f + a:15 * f + b] = False
    x_init = hb["states"].reshape(-1).copy()

    def fun(z):
        x = x_init.copy()
        x[free] = z
        return residual_vector(oracle, P, hb, x)

    sol = least_squares(fun, states.reshape(-1)[free], method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=200)
    c_sp = 0.5 * np.sum(sol.fun ** 2)
    # scipy started from the oracle's answer may still polish it; both costs must agree closely and the oracle's
    # solution must already be (numerically) stationary
    # The Ceres-style LM creeps along the kink of the norm-type ground/wheel residuals, so after 200
    # iterations it sits within 1e-4 (relative cost) of the stationary point scipy polishes it to.
    print("oracle cost", c_or, "scipy cost", c_sp, "max |dx|", np.abs(sol.x - states.reshape(-1)[free]).max())
    assert c_sp <= c_or * (1 + 1e-9)
    assert abs(c_or - c_sp) <= 1e-4 * max(c_sp, 1.0)
    d = np.abs(sol.x - states.reshape(-1)[free])
    assert d.max() < 2e-3
    assert summ["final_cost"][0] <= summ["initial_cost"][0]


def test_lm_iteration_cap_and_summary(oracle):
    P = L.corridor_params(max_iters=3)
    sb, hb = small_window(oracle, P)
    st, summ = oracle.solve(P, hb)
    assert summ["iterations"][0] == 3 and summ["termination"][0] == abi.TERM_NO_CONVERGENCE
    assert summ["num_successful_steps"][0] + summ["num_unsuccessful_steps"][0] == 3
    # constant blocks never move (solver.cpp:787-794)
    assert np.array_equal(st[0, 0:6], hb["states"].reshape(-1, 15)[0, 0:6])


def test_marginalisation_against_dense_numpy(oracle):
    P = L.corridor_params(max_iters=10)
    for mk in (lambda: L.synth.config_tracking2(1), lambda: L.synth.config_init(1, n_frames=4)):
        sb = mk()
        hb = oracle.preintegrate_batch(P, sb)
        H, g, _ = oracle.linearize(P, hb, mode=1)
        X0, J, r, dH, dg = oracle.marginalize(P, hb)
        m = H.shape[1] - 15
        Hmm, Hmr, Hrr = H[0, :m, :m], H[0, :m, m:], H[0, m:, m:]
        gg = -g[0]
        want_H = Hrr - Hmr.T @ np.linalg.solve(Hmm, Hmr)
        want_g = gg[m:] - Hmr.T @ np.linalg.solve(Hmm, gg[:m])
        np.testing.assert_allclose(dH[0], want_H, rtol=1e-6, atol=1e-7 * np.abs(want_H).max())
        np.testing.assert_allclose(dg[0], want_g, rtol=1e-6, atol=1e-7 * np.abs(want_g).max())
        # J_lin^T J_lin reproduces the Schur complement on the eigen-space kept (> 1e-8, solver.cpp:390-397)
        w, V = np.linalg.eigh(0.5 * (want_H + want_H.T))
        keep = w > 1e-8
        np.testing.assert_allclose(J[0].T @ J[0], (V[:, keep] * w[keep]) @ V[:, keep].T, rtol=1e-6, atol=1e-7 * np.abs(want_H).max())
        np.testing.assert_allclose(J[0].T @ r[0], -(V[:, keep] @ V[:, keep].T) @ want_g, rtol=1e-5, atol=1e-6 * np.abs(want_g).max())
        np.testing.assert_array_equal(X0[0], hb["states"].reshape(-1, 15)[-1])
