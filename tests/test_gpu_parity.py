"""GPU parity tests: the CUDA path, called through the C ABI (include/lvio2d.h), against the CPU oracle on the same
seeded inputs.  Tolerances (all float64): factor values/Jacobians 1e-9 relative; normal equations 1e-9 relative to
the largest entry; solved poses 1e-6 m / rad (the north-star bar is 1e-4 m / 1e-4 rad per keyframe)."""
import ctypes as C

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200 import abi
from lvio2d_b200.params import params_T

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(11)


@pytest.fixture(scope="module")
def P10():
    return L.corridor_params(max_iters=10)


@pytest.fixture(scope="module", params=[256, 128, 32], ids=["8warps", "4warps", "1warp"])
def ctx(P10, request):
    """One context per thread-group shape of window_kernel (LVIO2D_WINDOW_THREADS is read by lvio2d_create)."""
    import os

    from lvio2d_b200.solver import Context

    old = os.environ.get("LVIO2D_WINDOW_THREADS")
    os.environ["LVIO2D_WINDOW_THREADS"] = str(request.param)
    c = Context(P10)
    if old is None:
        del os.environ["LVIO2D_WINDOW_THREADS"]
    else:
        os.environ["LVIO2D_WINDOW_THREADS"] = old
    yield c
    c.close()


def relclose(a, b, rtol, name=""):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{name}: rel err {err:.3e} > {rtol}"


def pose_like(sb, f):
    return sb.states[f, 0:6].copy()


def test_library_is_cuda_only(P10):
    from lvio2d_b200 import solver

    lib = solver.load_library()
    for name in solver.EXPORTS:
        assert hasattr(lib, name)


def test_factor_hooks_match_oracle(ctx, oracle, P10):
    sb = L.synth.make_batch(1, 7, n_frames=4, beams=90)
    hb = oracle.preintegrate_batch(P10, sb)
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)
    for i in range(1, 4):
        si, sj = sb.states[i - 1], sb.states[i]
        r, J = ctx.eval_imu_factor(imu[i - 1], si, sj)
        orr, oJ = oracle.eval_imu_factor(P10, imu[i - 1], si, sj)
        relclose(r, orr, 1e-9, "imu res")
        relclose(J, oJ, 1e-9, "imu jac")
        r, J = ctx.eval_wheel_factor(wheel[i - 1], si[:6], sj[:6])
        orr, oJ = oracle.eval_wheel_factor(P10, wheel[i - 1], si[:6], sj[:6])
        relclose(r, orr, 1e-9, "wheel res")
        relclose(J, oJ, 1e-8, "wheel jac")
        r, J = ctx.eval_ground_factors(sj[:6])
        orr, oJ = oracle.eval_ground_factors(P10, sj[:6])
        relclose(r, orr, 1e-9, "ground res")
        relclose(J, oJ, 1e-9, "ground jac")
    for _ in range(10):
        pi_, pj_ = sb.states[0, :6], sb.states[2, :6]
        l1_p1, l1_p2 = np.append(RNG.uniform(-6, 6, 2), 0), np.append(RNG.uniform(-6, 6, 2), 0)
        l2_p1, l2_p2 = np.append(RNG.uniform(-6, 6, 2), 0), np.append(RNG.uniform(-6, 6, 2), 0)
        r, J = ctx.eval_laser_factor(l1_p1, l1_p2, l2_p1, l2_p2, pi_, pj_)
        orr, oJ = oracle.eval_laser_factor(P10, l1_p1, l1_p2, l2_p1, l2_p2, pi_, pj_)
        relclose(r, orr, 1e-10, "laser res")
        relclose(J, oJ, 1e-9, "laser jac")


def test_preintegration_matches_oracle(ctx, oracle, P10):
    sb = L.synth.make_batch(2, 3, n_frames=6, beams=32)
    got = ctx.imu_preintegrate(sb.imu_offset, sb.imu_samples, sb.bias0)
    want = oracle.imu_preintegrate(P10, sb.imu_offset, sb.imu_samples, sb.bias0)
    relclose(got[:, :15], want[:, :15], 1e-11, "imu X")
    relclose(got[:, 15:240], want[:, 15:240], 1e-11, "imu J")
    relclose(got[:, 240:465], want[:, 240:465], 1e-7, "imu sqrtP")  # cond(P) ~ 1e8 amplifies the inverse's rounding
    relclose(got[:, 465], want[:, 465], 1e-15, "imu Dt")
    gw = ctx.wheel_preintegrate(sb.wheel_offset, sb.wheel_steps)
    ww = oracle.wheel_preintegrate(P10, sb.wheel_offset, sb.wheel_steps)
    relclose(gw, ww, 1e-12, "wheel blob")
    # empty interval and ignored steps
    blob = ctx.wheel_preintegrate([0, 0, 2], np.array([[12.0, 1, 1, 1, 1, 1, 1], [0.05, 0.5, 0, 0, 0, 0, 0.1]]))
    ob = oracle.wheel_preintegrate(P10, [0, 0, 2], np.array([[12.0, 1, 1, 1, 1, 1, 1], [0.05, 0.5, 0, 0, 0, 0, 0.1]]))
    relclose(blob, ob, 1e-12, "wheel edge")


CASES = {
    "c1": lambda: L.synth.config_c1(),
    "c2_small": lambda: L.synth.make_batch(2, 42, n_frames=5, beams=300, fov_deg=270.0),
    "tracking2": lambda: L.synth.config_tracking2(2),
    "init": lambda: L.synth.config_init(2, n_frames=6),
    "init_beam": lambda: L.synth.config_init(1, n_frames=4, mode="beam"),
    "c2_full": lambda: L.synth.config_c2(1),
}


@pytest.mark.parametrize("case", ["c1", "c2_small", "tracking2", "init", "init_beam"])
@pytest.mark.parametrize("mode", [0, 1])
def test_linearize_matches_oracle(ctx, oracle, P10, case, mode):
    sb = CASES[case]()
    hb = oracle.preintegrate_batch(P10, sb)
    ctx.set_windows(hb)
    H, g, cost = ctx.linearize(mode)
    oH, og, ocost = oracle.linearize(P10, hb, mode=mode)
    relclose(cost, ocost, 1e-11, "cost")
    relclose(g, og, 1e-9, "gradient")
    relclose(H, oH, 1e-9, "hessian")
    assert np.allclose(H, np.swapaxes(H, 1, 2), rtol=0, atol=1e-9 * np.abs(H).max())


@pytest.mark.parametrize("case,iters", [("c1", 1), ("c2_small", 10), ("tracking2", 20), ("init", 20), ("c2_full", 10),
                                        ("tracking2", 50), ("init", 50)])
@pytest.mark.parametrize("wt", [32, 128, 256, 512])
def test_solve_matches_oracle(oracle, case, iters, wt, monkeypatch):
    """Up to ~20 iterations the two minimisers walk in lock-step (differences ~1e-13).  The reference's default of 50
    iterations (solver.cpp:161-168, no fast_mode) ends in a slowly converging zig-zag (the ground/wheel residuals are
    norms) where rounding differences are amplified and an accept/reject decision can flip; there
    the bar is the north-star's 1e-4 m / 1e-4 rad per keyframe, and costs within 1 %."""
    from lvio2d_b200.solver import Context

    # both thread-group shapes of window_kernel: one warp per window (the batched shape) and four warps per window (what
    # small batches get by default)
    monkeypatch.setenv("LVIO2D_WINDOW_THREADS", str(wt))
    if wt == 512:
        # sixteen warps per window + cyclic reduction over the frames (tracking topology only; the initialisation
        # topology keeps eight warps): the three-kernel loop, not the fused small-batch kernel (which runs eight warps)
        monkeypatch.setenv("LVIO2D_FUSED_SMALL", "0")
    P = L.corridor_params(max_iters=iters)
    sb = CASES[case]()
    hb = oracle.preintegrate_batch(P, sb)
    with Context(P) as c:
        c.set_windows(hb)
        summ = c.solve()
        got = c.get_states()
    want, osumm = oracle.solve(P, hb)
    print(case, "gpu", summ, "oracle", osumm)
    d = np.abs(got - want)
    relclose(summ["initial_cost"], osumm["initial_cost"], 1e-11, "initial cost")
    assert np.all(summ["final_cost"] <= summ["initial_cost"])
    if iters <= 20:
        assert np.array_equal(summ["iterations"], osumm["iterations"])
        assert np.array_equal(summ["termination"], osumm["termination"])
        assert np.array_equal(summ["num_successful_steps"], osumm["num_successful_steps"])
        relclose(summ["final_cost"], osumm["final_cost"], 1e-9, "final cost")
        relclose(summ["final_radius"], osumm["final_radius"], 1e-6, "final radius")
        assert d[:, 0:6].max() < 1e-9 and d[:, 6:].max() < 1e-9, (d[:, 0:6].max(), d[:, 6:].max())
    elif case == "init":
        # the initialisation problem stops at the 50-iteration cap deep inside the zig-zag regime: two CPU implementations
        # of the same reference text (the oracle and solver.cpp compiled against the stub Eigen / Ceres tree) already
        # differ by 9.4e-5 there (tests/test_ref_solver.py::test_init_solve_matches_reference_text measures it and compares
        # in lock-step under a 15-iteration cap instead); an equally good answer is required
        relclose(summ["final_cost"], osumm["final_cost"], 5e-2, "final cost")
        assert d[:, 0:3].max() < 1e-3 and d[:, 3:6].max() < 1e-3, (d[:, 0:3].max(), d[:, 3:6].max())
    else:
        relclose(summ["final_cost"], osumm["final_cost"], 1e-2, "final cost")
        assert d[:, 0:3].max() < 1e-4 and d[:, 3:6].max() < 1e-4, (d[:, 0:3].max(), d[:, 3:6].max())


def test_marginalize_matches_oracle(ctx, oracle, P10):
    for case in ("tracking2", "init"):
        sb = CASES[case]()
        hb = oracle.preintegrate_batch(P10, sb)
        ctx.set_windows(hb)
        X0, J, r = ctx.marginalize()
        oX0, oJ, orr, odH, odg = oracle.marginalize(P10, hb)
        relclose(X0, oX0, 1e-15, "X0")
        # the prior only enters through J^T J (marginalization_factor.h:50 drops linearized_R); rows are defined
        # up to sign / rotation inside (near-)degenerate eigenspaces
        JTJ, oJTJ = np.einsum("bki,bkj->bij", J, J), np.einsum("bki,bkj->bij", oJ, oJ)
        relclose(JTJ, oJTJ, 1e-7, "J_lin^T J_lin")
        relclose(np.abs(np.linalg.norm(J, axis=2)), np.abs(np.linalg.norm(oJ, axis=2)), 1e-6, "row norms (sqrt eigenvalues)")
        relclose(np.einsum("bki,bk->bi", J, r), np.einsum("bki,bk->bi", oJ, orr), 1e-6, "J^T r")


def test_solver_class_mirrors_reference_flow(oracle, P10):
    """lvio_2d::solver surface: solve -> marginalization -> solve with the prior, like trajectory::do_tracking."""
    from lvio2d_b200.solver import FrameInfo, LaserMatch, Line, Solver

    P = L.corridor_params(max_iters=50)
    sb = L.synth.config_tracking2(1)
    hb = oracle.preintegrate_batch(P, sb)
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)
    frames = []
    for i in range(2):
        s = sb.states[i]
        f = FrameInfo(0.1 * i, s[0:3], s[3:6], s[6:9], s[9:15], imu[i - 1] if i else None, wheel[i - 1] if i else None)
        a, b = int(sb.point_offset[i]), int(sb.point_offset[i + 1])
        l0 = int(sb.line_offset[i])
        lines1, lines2 = [], []
        for k in range(a, b, 2):
            li = sb.lines[l0 + sb.point_line[k]]
            lines1.append(Line([li[0], li[1], 0], [li[2], li[3], 0]))
            lines2.append(Line([*sb.points[k], 0], [*sb.points[k + 1], 0]))
        f.add_laser_match(LaserMatch(lines1, lines2, sb.ref_pose[i, 0:3], sb.ref_pose[i, 3:6]))
        frames.append(f)
    sol = Solver(P, fast_mode=False)
    # the oracle solves the very batch the Solver flattens the frames into (no prior yet, frame 0 constant)
    hb_s = sol._batch(frames, "tracking", with_prior=True, laser_frames={1})
    assert hb_s.prior_frame == -1 and hb_s["const_mask"][0] == (abi.CONST_P | abi.CONST_Q)
    P20 = L.corridor_params(max_iters=20)
    sol.ctx.close()
    sol = Solver(P20, fast_mode=False)
    want, _ = oracle.solve(P20, hb_s)
    sol.solve(frames)
    got = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in frames])
    assert np.abs(got - want).max() < 1e-9
    assert np.array_equal(frames[-1].laser_match.p2, frames[-1].p)
    sol.marginalization(frames)
    assert sol.has_linearized_block and sol.linearized_jacobians.shape == (15, 15)
    hb_m = sol._batch(frames, "marg", with_prior=False, laser_frames={0, 1})
    oX0, oJ, _, _, _ = oracle.marginalize(P20, hb_m)
    relclose(sol.linearized_jacobians.T @ sol.linearized_jacobians, oJ[0].T @ oJ[0], 1e-6, "prior information")
    assert np.array_equal(sol.linearized_X, got[1])
    # second solve of the flow: the prior now sits on frame n-2
    hb2 = sol._batch(frames, "tracking", with_prior=True, laser_frames={1})
    assert hb2.prior_frame == 0
    sol.solve(frames)
    want2, _ = oracle.solve(P20, hb2)
    got2 = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in frames])
    assert np.abs(got2 - want2).max() < 1e-8
    sol.close()


def test_point_sharded_solve_matches_unsharded(oracle, P10):
    """Two ranks, each owning half of every frame's points, exchanging only the summed per-frame blocks once per
    iteration (what the NCCL all-reduce does in bench.py --gpus N), reproduce the unsharded solve."""
    import torch
    from lvio2d_b200.solver import Context

    sb = L.synth.make_batch(2, 5, n_frames=4, beams=257)
    hb = oracle.preintegrate_batch(P10, sb)
    with Context(P10) as ref:
        ref.set_windows(hb)
        ref.solve()
        want = ref.get_states()
    ranks = []
    for rank in range(2):
        c = Context(P10)
        c.set_windows(hb)
        c.set_point_shard(rank, 2)
        _, n = c.reduce_buffer()
        t = torch.zeros(n, dtype=torch.float64, device="cuda")
        c.set_reduce_buffer(t.data_ptr(), n)
        c.solve_begin()
        ranks.append((c, t))
    for it in range(P10.max_iters + 1):
        for c, _ in ranks:
            c.eval_laser()
            c.sync()
        total = ranks[0][1] + ranks[1][1]
        for _, t in ranks:
            t.copy_(total)
        torch.cuda.synchronize()
        active = [c.lm_step(want_active=True) for c, _ in ranks]
        assert active[0] == active[1]
    for c, _ in ranks:
        got = c.get_states()
        assert np.abs(got - want).max() < 1e-9
        c.close()


@pytest.mark.parametrize("case", ["c2_small", "init_beam", "tracking2"])
@pytest.mark.parametrize("huber,assoc", [(1.5, 0), (0.0, 1), (1.5, 1)])
def test_huber_and_in_kernel_association_match_oracle(oracle, case, huber, assoc):
    """The two extensions the north-star names on top of the reference (which has neither: solver.cpp:635,
    trajectory.cpp:210): Ceres-style Huber loss on the laser residuals and nearest-line re-association at every
    evaluation (BASELINE config 3).  Parity is against the oracle running the same rules."""
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=8)
    P.huber_delta = huber
    P.assoc_mode = assoc
    sb = CASES[case]()
    hb = oracle.preintegrate_batch(P, sb)
    with Context(P) as c:
        c.set_windows(hb)
        for mode in (0, 1):
            H, g, cost = c.linearize(mode)
            oH, og, ocost = oracle.linearize(P, hb, mode=mode)
            relclose(cost, ocost, 1e-10, "cost")
            relclose(g, og, 1e-8, "gradient")
            relclose(H, oH, 1e-8, "hessian")
        summ = c.solve()
        got = c.get_states()
    want, osumm = oracle.solve(P, hb)
    assert np.array_equal(summ["num_successful_steps"], osumm["num_successful_steps"])
    relclose(summ["final_cost"], osumm["final_cost"], 1e-8, "final cost")
    assert np.abs(got - want).max() < 1e-8
    if huber > 0 and not assoc:
        # the loss must actually bite on the initial point of these cases
        P0 = L.corridor_params(max_iters=8)
        assert oracle.cost(P, hb)[0] < oracle.cost(P0, hb)[0]


def test_out_of_range_correspondences(oracle):
    """A point_line index beyond its frame's line list: rejected on the host-buffer path (LVIO2D_ERR_INVALID_ARG); on a path
    the host cannot scan (here: the wire encoding) the kernel skips the point instead of reading outside its line table."""
    from lvio2d_b200 import abi
    from lvio2d_b200.solver import Context, Lvio2dError

    P = L.corridor_params(max_iters=4)
    hb = oracle.preintegrate_batch(P, L.synth.make_batch(1, 5, n_frames=4, beams=361, fov_deg=270.0))
    pl = hb["point_line"].copy()
    nl = int(np.diff(hb["line_offset"])[1])
    k = int(hb["point_offset"][1]) + 3
    pl[k] = nl + 7
    with Context(P) as c:
        with pytest.raises(Lvio2dError) as e:
            c.set_windows(hb.replace(point_line=pl))
        assert e.value.status == abi.ERR_INVALID_ARG
        # the same batch with that point dropped is what the guarded kernel must compute
        drop = pl.copy()
        drop[k] = -1
        c.set_windows(hb.replace(point_line=drop))
        H0, g0, c0 = c.linearize(0)
    if np.all(np.diff(hb["point_offset"]) == 361):
        import math
        wire = abi.ScanWire.from_points(hb, 361, np.float32(math.radians(-135.0)), np.float32(math.radians(270.0) / 360))
        wire.beam_line.reshape(-1)[k] = nl + 7
        pts, line, off = wire.points()
        line[k] = -1
        with Context(P) as c:
            c.set_windows_wire(hb.replace(points=None, point_line=None, point_offset=None), wire)
            H1, g1, c1 = c.linearize(0)
            c.set_windows(hb.replace(points=pts, point_line=line, point_offset=off))
            H2, g2, c2 = c.linearize(0)
        # (the device's sincos and numpy's cos / sin differ in the last bit of a few points)
        assert np.abs(H1 - H2).max() <= 1e-12 * np.abs(H2).max() and np.abs(g1 - g2).max() <= 1e-11 * np.abs(g2).max() and np.abs(c1 - c2).max() <= 1e-12 * np.abs(c2).max()


def test_association_grid_equals_the_all_lines_loop(oracle, monkeypatch):
    """BASELINE config 3: the precomputed candidate grid (16 x 16 cells over the local map, one bit per line) must pick
    exactly the line the all-lines loop picks — bit-identical normal equations and states, with ragged line counts, points
    outside the map and a gate / cut-off other than the defaults."""
    from lvio2d_b200.solver import Context

    for gate, md in ((0.0, 0.0), (0.03, 0.2)):
        P = L.corridor_params(max_iters=6)
        P.assoc_mode = 1
        P.assoc_gate, P.assoc_max_dist = gate, md
        sb = L.synth.make_batch(3, 11, n_frames=6, beams=400, fov_deg=270.0)
        hb = oracle.preintegrate_batch(P, sb)
        pts = hb["points"].reshape(-1, 2).copy()
        pts[::37] *= 40.0                                 # a few points far outside the map
        hb = hb.replace(points=pts)
        res = {}
        for grid in ("1", "0"):
            monkeypatch.setenv("LVIO2D_ASSOC_GRID", grid)
            with Context(P) as c:
                c.set_windows(hb)
                H, g, cost = c.linearize(0)
                c.solve()
                res[grid] = (H, g, cost, c.get_states())
        for a, b in zip(res["1"], res["0"]):
            assert np.array_equal(a, b)
        want, _ = oracle.solve(P, hb)
        assert np.abs(res["1"][3] - want).max() < 1e-8


def test_paired_and_single_factor_kernels_agree(oracle, monkeypatch):
    """factor_pair_kernel (two items per warp, 12 dual columns + 15 closed-form columns; the default) against factor_kernel
    (one item per warp, all 30 columns in dual arithmetic): same normal equations, same solve."""
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=10)
    res = {}
    for case in ("c2_small", "tracking2", "init"):
        hb = oracle.preintegrate_batch(P, CASES[case]())
        for paired in ("1", "0", "auto"):
            if paired == "auto":
                monkeypatch.delenv("LVIO2D_FACTOR_PAIRED")     # the library's own rule: one item per warp for a batch this small
            else:
                monkeypatch.setenv("LVIO2D_FACTOR_PAIRED", paired)
            with Context(P) as c:
                c.set_windows(hb)
                H, g, cost = c.linearize(0)
                c.solve()
                res[paired] = (H, g, cost, c.get_states())
        assert all(np.array_equal(x, y) for x, y in zip(res["auto"], res["0"]))
        for a, b in zip(res["1"], res["0"]):
            relclose(a, b, 1e-12, case)


def test_fused_small_solve_matches_three_kernel_loop(oracle, monkeypatch):
    """solve_small_kernel (whole minimiser loop in one launch, the default for batches of small windows) against the
    three-kernel loop on the reference's own problem shapes: identical iteration counts and states to 1e-12."""
    from lvio2d_b200.solver import Context

    for case, iters in (("tracking2", 20), ("init", 20), ("c2_small", 10)):
        P = L.corridor_params(max_iters=iters)
        hb = oracle.preintegrate_batch(P, CASES[case]())
        res = {}
        # the fused kernel runs eight warps per window; the loop is pinned to the same thread-group shape (its default for a
        # batch this small is the cyclic-reduction shape, which eliminates in another order: equal to ~1e-12, covered by
        # test_solve_matches_oracle[512-*])
        monkeypatch.setenv("LVIO2D_WINDOW_THREADS", "256")
        for fused in ("1", "0"):
            monkeypatch.setenv("LVIO2D_FUSED_SMALL", fused)
            with Context(P) as c:
                c.set_windows(hb)
                summ = c.solve()
                res[fused] = (summ["iterations"].copy(), summ["final_cost"].copy(), c.get_states())
        assert np.array_equal(res["1"][0], res["0"][0]), case
        relclose(res["1"][1], res["0"][1], 1e-12, case + " cost")
        relclose(res["1"][2], res["0"][2], 1e-12, case + " states")


def _ragged(hb, keep_points, keep_lines=None, unmatched_frame=None):
    """The same windows with frame f keeping only its first keep_points[f] points (and keep_lines[f] lines; points whose
    line was cut become unmatched), optionally every point of one frame unmatched."""
    po, lo = hb["point_offset"], hb["line_offset"]
    pts, pl, ln = hb["points"].reshape(-1, 2), hb["point_line"].reshape(-1), hb["lines"].reshape(-1, 4)
    F = len(po) - 1
    P2, L2, I2, O2, LO2 = [], [], [], [0], [0]
    for f in range(F):
        kp = min(int(keep_points[f]), int(po[f + 1] - po[f]))
        kl = int(lo[f + 1] - lo[f]) if keep_lines is None else min(int(keep_lines[f]), int(lo[f + 1] - lo[f]))
        idx = pl[po[f]:po[f] + kp].copy()
        idx[idx >= kl] = -1
        if unmatched_frame is not None and f == unmatched_frame:
            idx[:] = -1
        P2.append(pts[po[f]:po[f] + kp]); I2.append(idx); L2.append(ln[lo[f]:lo[f] + kl])
        O2.append(O2[-1] + kp); LO2.append(LO2[-1] + kl)
    return hb.replace(points=np.concatenate(P2).reshape(-1, 2), point_line=np.concatenate(I2), point_offset=np.array(O2, np.int64),
                      lines=np.concatenate(L2).reshape(-1, 4), line_offset=np.array(LO2, np.int64), point_weight=None)


@pytest.mark.parametrize("wt", [32, 256, 512])
def test_ragged_and_degenerate_windows_match_oracle(oracle, wt, monkeypatch):
    """Ragged batches take the offset-table path of the scan-match kernel (the bench's uniform batches take the arithmetic
    one): frames with different point and line counts, a frame without points, a frame whose points are all unmatched, a
    frame whose local map lost most of its lines — normal equations and the 10-iteration solve against the oracle."""
    from lvio2d_b200.solver import Context

    monkeypatch.setenv("LVIO2D_WINDOW_THREADS", str(wt))
    monkeypatch.setenv("LVIO2D_FUSED_SMALL", "0")
    P = L.corridor_params(max_iters=10)
    sb = L.synth.make_batch(3, 77, n_frames=6, beams=240, fov_deg=270.0)
    hb0 = oracle.preintegrate_batch(P, sb)
    F = 18
    g = np.random.default_rng(5)
    keep = g.integers(40, 241, F)
    keep[4] = 0                                  # a frame without points
    keep[9] = 1
    lines = g.integers(3, 60, F)
    lines[13] = 2                                # most correspondences of this frame point at lines that are gone
    hb = _ragged(hb0, keep, lines, unmatched_frame=7)
    assert len(set(np.diff(hb["point_offset"]))) > 5
    with Context(P) as c:
        c.set_windows(hb)
        H, gr, cost = c.linearize(0)
        summ = c.solve()
        got = c.get_states()
    oH, og, ocost = oracle.linearize(P, hb, mode=0)
    relclose(cost, ocost, 1e-11, "cost")
    relclose(gr, og, 1e-9, "gradient")
    relclose(H, oH, 1e-9, "hessian")
    want, osumm = oracle.solve(P, hb)
    assert np.array_equal(summ["iterations"], osumm["iterations"]) and np.array_equal(summ["num_successful_steps"], osumm["num_successful_steps"])
    assert np.abs(got - want).max() < 1e-9
    assert np.all(summ["final_cost"] <= summ["initial_cost"])


def test_longest_window_and_single_frame(oracle):
    """n_frames = 64 is the longest window the ABI accepts (65 is a domain error); a window of one frame has no IMU /
    wheel factor at all."""
    from lvio2d_b200.solver import Context, Lvio2dError

    P = L.corridor_params(max_iters=5)
    for n in (64, 1):
        sb = L.synth.make_batch(2, 11, n_frames=n, beams=60, fov_deg=270.0)
        hb = oracle.preintegrate_batch(P, sb)
        with Context(P) as c:
            c.set_windows(hb)
            summ = c.solve()
            got = c.get_states()
        want, osumm = oracle.solve(P, hb)
        assert np.array_equal(summ["iterations"], osumm["iterations"]), n
        assert np.abs(got - want).max() < 1e-8, (n, np.abs(got - want).max())
    sb = L.synth.make_batch(1, 11, n_frames=65, beams=20, fov_deg=270.0)
    hb = oracle.preintegrate_batch(P, sb)
    with Context(P) as c:
        with pytest.raises(Lvio2dError):
            c.set_windows(hb)


def test_bench_size_batch_properties():
    """At the bench's size (148 x 8 windows of 30 x 1081, no oracle run possible): identical windows give identical
    results wherever they sit in the batch, costs never increase, every window reports the iteration cap, and a second
    solve started from the first one's result only descends further."""
    import torch

    import bench
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=bench.MAX_ITERS)
    dev = torch.device("cuda:0")
    with Context(P) as c:
        hb, uniq = bench.build_host_batch(c, 1184, seed0=42, config="c2")
        d, keep = bench.to_device_struct(hb, torch, dev)
        c.bind_windows(d, keepalive=keep)
        summ = c.solve()
        x = c.get_states().reshape(1184, -1)
        assert np.all(summ["final_cost"] <= summ["initial_cost"]) and np.all(np.isfinite(x))
        assert np.all(summ["iterations"] == bench.MAX_ITERS)
        for w in range(uniq, 1184):          # window w is a copy of window w % uniq
            assert np.array_equal(x[w], x[w % uniq]), w
        c.reset_states(x.reshape(-1, 15))
        summ2 = c.solve()
        x2 = c.get_states().reshape(1184, -1)
        assert np.all(summ2["final_cost"] <= summ["final_cost"] * (1 + 1e-12))
        moved = np.abs(x2 - x).reshape(1184, 30, 15)
        assert moved[:, :, 0:6].max() < 0.1           # ten more iterations from the 10-iteration point (the zig-zag directions still move by centimetres)
