"""Size-independent properties at BASELINE.json's full sizes (the oracle needs seconds per window there, so these
tests lean on invariants instead of a second implementation)."""
import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2_batch(oracle):
    P = L.corridor_params(max_iters=10)
    sb = L.synth.config_c2(3)
    return P, sb, oracle.preintegrate_batch(P, sb)


def test_replicated_windows_give_identical_results(c2_batch):
    """Every window of a batch is solved independently: tiling a batch must reproduce the same states bit for bit,
    whatever the grid / tile configuration the larger batch selects."""
    from lvio2d_b200.solver import Context

    P, sb, hb = c2_batch
    with Context(P) as c:
        c.set_windows(hb)
        c.solve()
        base = c.get_states().reshape(hb.n_windows, -1)
    big = L.synth.tile_batch(hb, 40)   # 120 windows: different tiles-per-frame than 3 windows
    with Context(P) as c:
        c.set_windows(big)
        summ = c.solve()
        st = c.get_states().reshape(big.n_windows, -1)
    for k in range(big.n_windows):
        assert np.abs(st[k] - base[k % hb.n_windows]).max() < 1e-10
    assert np.all(summ["final_cost"] <= summ["initial_cost"])
    assert np.all(summ["iterations"] == 10)


def test_full_c2_window_converges_towards_truth_and_matches_oracle_cost(oracle, c2_batch):
    from lvio2d_b200.solver import Context

    P, sb, hb = c2_batch
    with Context(P) as c:
        c.set_windows(hb)
        summ = c.solve()
        st = c.get_states()
    # the solved states, re-evaluated by the oracle's cost function, reproduce the device's final cost
    oc = oracle.cost(P, hb, st)
    np.testing.assert_allclose(summ["final_cost"], oc, rtol=1e-9)
    err0 = np.abs(sb.states - sb.truth)[:, 0:2].max()
    err1 = np.abs(st - sb.truth)[:, 0:2].max()
    assert err1 < 0.25 * err0, (err0, err1)
    # constant blocks (frame 0 pose) did not move
    n = hb.n_frames
    assert np.array_equal(st[::n, 0:6], sb.states[::n, 0:6])


def test_idempotence_of_a_converged_point(c2_batch):
    """Solving again from the solution: the cost cannot go up and the states barely move."""
    from lvio2d_b200.solver import Context

    P, sb, hb = c2_batch
    P50 = L.corridor_params(max_iters=50)
    with Context(P50) as c:
        c.set_windows(hb)
        s1 = c.solve()
        x1 = c.get_states()
        c.set_windows(hb.replace(states=x1))
        s2 = c.solve()
        x2 = c.get_states()
    assert np.all(s2["final_cost"] <= s1["final_cost"] * (1 + 1e-12))
    assert np.abs(x2 - x1)[:, 0:6].max() < 5e-3


def test_c4_point_sharding_matches_unsharded(oracle):
    """BASELINE config 4 (4096 beams x 50 frames): 4 ranks' slices, blocks summed once per iteration."""
    import torch
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=5)
    sb = L.synth.config_c4(1)
    hb = oracle.preintegrate_batch(P, sb)
    assert hb.n_points > 150000
    with Context(P) as ref:
        ref.set_windows(hb)
        ref.solve()
        want = ref.get_states()
    world = 4
    ranks = []
    for rank in range(world):
        c = Context(P)
        c.set_windows(hb)
        c.set_point_shard(rank, world)
        _, n = c.reduce_buffer()
        t = torch.zeros(n, dtype=torch.float64, device="cuda")
        c.set_reduce_buffer(t.data_ptr(), n)
        c.solve_begin()
        ranks.append((c, t))
    for it in range(P.max_iters + 1):
        for c, _ in ranks:
            c.eval_laser()
            c.sync()
        total = sum(t for _, t in ranks)
        for _, t in ranks:
            t.copy_(total)
        torch.cuda.synchronize()
        for c, _ in ranks:
            c.lm_step()
    for c, _ in ranks:
        got = c.get_states()
        assert np.abs(got - want).max() < 1e-8
        c.close()


def test_error_paths(oracle):
    from lvio2d_b200.solver import Context, Lvio2dError

    P = L.corridor_params(max_iters=2)
    with Context(P) as c:
        with pytest.raises(Lvio2dError) as e:
            c.lib.lvio2d_solve.restype  # touch
            c._check(c.lib.lvio2d_solve(c._h, None), "lvio2d_solve")
        assert e.value.status == abi.ERR_NO_WINDOW
        sb = L.synth.config_init(1, n_frames=3)
        hb = oracle.preintegrate_batch(P, sb)
        bad = hb.replace(ref_frame=np.array([-1, 0, 1], np.int32))   # reference frame must be frame 0
        with pytest.raises(Lvio2dError) as e:
            c.set_windows(bad)
        assert e.value.status == abi.ERR_DOMAIN
        # empty laser input: a window with only IMU / wheel / ground / prior factors still solves
        sb2 = L.synth.make_batch(1, 3, n_frames=3, beams=20)
        hb2 = oracle.preintegrate_batch(P, sb2)
        nolaser = hb2.replace(point_offset=np.zeros(4, np.int64), points=np.zeros((0, 2)), point_line=np.zeros(0, np.int32),
                              line_offset=np.zeros(4, np.int64), lines=np.zeros((0, 4)))
        c.set_windows(nolaser)
        summ = c.solve()
        want, osumm = oracle.solve(P, nolaser)
        assert np.abs(c.get_states() - want).max() < 1e-9
    bad_params = L.corridor_params()
    bad_params.abi_version = 99
    with pytest.raises(Lvio2dError):
        Context(bad_params)
