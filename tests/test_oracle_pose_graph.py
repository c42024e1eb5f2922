"""Oracle of the back-end pose-graph solve (oracle/pose_graph.hpp = keyframe_manager::solve + edge_factor, reference
src/trajectory/keyframe_manager.cpp:722-838, src/factor/edge_factor.h:4-27, :79-126) — SURVEY.md section 8f rank 4, the next
row; there is no device path for it yet.  Pins: an independent scipy restatement of the residual, finite differences for
the Jacobian, scipy.optimize.least_squares for the converged minimiser."""
import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

import lvio2d_b200 as L
from lvio2d_b200.params import params_T


def edge_noise_J(sigma_p=(0.1, 0.1, 0.1), sigma_q=(0.01, 0.01, 0.01)):
    """edge_noise::edge_noise (edge_factor.h:14-26) as written, including J(1,2) where J(1,1) was meant."""
    J = np.eye(6)
    J[0, 0] = 1.0 / sigma_p[0]
    J[1, 2] = 1.0 / sigma_p[1]
    J[2, 2] = 1.0 / sigma_p[2]
    J[3, 3], J[4, 4], J[5, 5] = 1.0 / sigma_q[0], 1.0 / sigma_q[1], 1.0 / sigma_q[2]
    return J


def T_of(pose):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = Rotation.from_rotvec(pose[3:6]).as_matrix(), pose[0:3]
    return T


def edge_residual(tf12, w, Jn, pi, pj):
    E = np.linalg.inv(T_of(pj)) @ T_of(pi) @ tf12
    return w * (Jn @ np.r_[E[:3, 3], Rotation.from_matrix(E[:3, :3]).as_rotvec()])


def ground_residuals(P, pose):
    T_io = np.eye(4)
    T_io[:3, :] = params_T(P, "T_imu_to_wheel")
    T = T_of(pose) @ T_io
    z = T[:3, 2]
    return np.array([T[2, 3] / P.manifold_p_sigma, np.arcsin(np.linalg.norm(np.cross(z, [0, 0, 1.0]))) / P.manifold_q_sigma])


def test_edge_factor_matches_scipy_and_finite_differences(oracle):
    g = np.random.default_rng(8)
    Jn = edge_noise_J()
    for _ in range(50):
        pi, pj = np.r_[g.uniform(-5, 5, 3), g.normal(0, 0.5, 3)], np.r_[g.uniform(-5, 5, 3), g.normal(0, 0.5, 3)]
        tf12 = np.linalg.inv(T_of(pi)) @ T_of(pj) @ T_of(np.r_[g.normal(0, 0.05, 3), g.normal(0, 0.02, 3)])
        w = g.uniform(0.5, 10.0)
        r, J = oracle.eval_edge_factor(tf12[:3, :], w, Jn, pi, pj)
        assert np.abs(r - edge_residual(tf12, w, Jn, pi, pj)).max() < 1e-9 * max(1.0, np.abs(r).max())
        x0 = np.r_[pi, pj]
        fd = np.zeros((6, 12))
        for c in range(12):
            h = 1e-6
            xp, xm = x0.copy(), x0.copy()
            xp[c] += h
            xm[c] -= h
            fd[:, c] = (edge_residual(tf12, w, Jn, xp[:6], xp[6:]) - edge_residual(tf12, w, Jn, xm[:6], xm[6:])) / (2 * h)
        assert np.abs(J - fd).max() < 1e-5 * max(1.0, np.abs(J).max())


def make_graph(K=24, seed=4):
    """A planar loop of K key frames: noisy sequential edges, one loop-closure edge (weight loop_edge_k = 10)."""
    g = np.random.default_rng(seed)
    T_io = params_T(L.corridor_params(), "T_imu_to_wheel")
    truth = np.zeros((K, 6))
    for k in range(K):
        # the wheel (base) frame moves in the ground plane; the IMU pose follows through T_imu_to_wheel, so the ground
        # factors are satisfied at the truth
        a = 2 * np.pi * k / K
        R_wb = Rotation.from_euler("z", a + np.pi / 2).as_matrix()
        R_wi, p_wi = L.synth._imu_pose_from_base(R_wb, np.array([3 * np.cos(a), 3 * np.sin(a), 0.0]), T_io)
        truth[k, 0:3], truth[k, 3:6] = p_wi, Rotation.from_matrix(R_wi).as_rotvec()
    edges, tfs, ws = [], [], []
    for k in range(K - 1):
        noise = T_of(np.r_[g.normal(0, 0.02, 3), g.normal(0, 0.005, 3)])
        edges.append((k, k + 1)); tfs.append((np.linalg.inv(T_of(truth[k])) @ T_of(truth[k + 1]) @ noise)[:3, :]); ws.append(1.0)
    edges.append((K - 1, 0)); tfs.append((np.linalg.inv(T_of(truth[K - 1])) @ T_of(truth[0]))[:3, :]); ws.append(10.0)
    init = truth.copy()
    T = T_of(truth[0])
    for k in range(K - 1):                      # dead reckoning along the noisy edges
        T = T @ np.vstack([tfs[k], [0, 0, 0, 1]])
        init[k + 1] = np.r_[T[:3, 3], Rotation.from_matrix(T[:3, :3]).as_rotvec()]
    return truth, init, np.array(edges, np.int32), np.array(tfs), np.array(ws)


def test_pose_graph_solve_agrees_with_scipy_least_squares(oracle):
    """ground_q = asin(|z x e_z|)/sigma has a kink exactly where it is satisfied (the reference's quirk, SURVEY appendix
    B), so with it the minimiser crawls (736 successful steps in 1000 and still moving) — the converged point is pinned on
    the smooth problem (edges + ground_p), the full problem through its cost at both ends."""
    P = L.corridor_params(max_iters=50)
    P.function_tolerance, P.parameter_tolerance = 1e-14, 1e-14      # run to the minimiser, not to Ceres' default stop
    truth, init, edges, tfs, ws = make_graph()
    Jn = edge_noise_J()
    K = len(init)

    def residuals(xfree, with_q):
        X = np.vstack([init[0], xfree.reshape(K - 1, 6)])                              # key frame 0 is constant
        out = [edge_residual(np.vstack([tfs[e], [0, 0, 0, 1]]), ws[e], Jn, X[i], X[j]) for e, (i, j) in enumerate(edges)]
        out += [ground_residuals(P, X[k])[:2 if with_q else 1] for k in range(K)]
        return np.concatenate(out)

    def jacobian(xfree, with_q):
        # edge blocks from the oracle's Jet Jacobian (pinned against finite differences above), ground rows by central
        # differences over the one pose they touch — scipy only needs a good Jacobian, the minimiser is pinned by the
        # independent residual
        X = np.vstack([init[0], xfree.reshape(K - 1, 6)])
        ng = 2 if with_q else 1
        J = np.zeros((6 * len(edges) + ng * K, 6 * (K - 1)))
        for e, (i, j) in enumerate(edges):
            _, Je = oracle.eval_edge_factor(tfs[e], ws[e], Jn, X[i], X[j])
            for node, cols in ((i, Je[:, 0:6]), (j, Je[:, 6:12])):
                if node > 0:
                    J[6 * e:6 * e + 6, 6 * (node - 1):6 * node] = cols
        for k in range(1, K):
            for c in range(6):
                xp, xm = X[k].copy(), X[k].copy()
                xp[c] += 1e-6
                xm[c] -= 1e-6
                r0 = 6 * len(edges) + ng * k
                J[r0:r0 + ng, 6 * (k - 1) + c] = (ground_residuals(P, xp)[:ng] - ground_residuals(P, xm)[:ng]) / 2e-6
        return J

    got, summ = oracle.pose_graph_solve(P, init, edges, tfs, ws, Jn, ground_p=True, ground_q=False)
    ref = least_squares(residuals, init[1:].ravel(), jac=jacobian, args=(False,), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-10, max_nfev=100)
    want = np.vstack([init[0], ref.x.reshape(K - 1, 6)])
    assert summ["termination"][0] == 1 and summ["iterations"][0] < 20
    assert np.array_equal(got[0], init[0])
    assert abs(summ["final_cost"][0] - ref.cost) <= 1e-8 * max(1.0, ref.cost)
    assert np.abs(got - want).max() < 1e-6
    # the loop closure pulled the drifted end of the chain back
    assert np.linalg.norm(got[-1, 0:3] - truth[-1, 0:3]) < np.linalg.norm(init[-1, 0:3] - truth[-1, 0:3])

    # the reference's full problem (both ground factors on every key frame, Ceres' 50 iterations)
    got, summ = oracle.pose_graph_solve(P, init, edges, tfs, ws, Jn)
    r0, r1 = residuals(init[1:].ravel(), True), residuals(got[1:].ravel(), True)
    assert abs(summ["initial_cost"][0] - 0.5 * r0 @ r0) <= 1e-9 * (0.5 * r0 @ r0)
    assert abs(summ["final_cost"][0] - 0.5 * r1 @ r1) <= 1e-9 * (0.5 * r1 @ r1)
    assert summ["final_cost"][0] < 1e-3 * summ["initial_cost"][0]
