"""Golden vectors (tests/golden/*.npz, written by scripts/make_golden.py with the oracle): the frozen solver inputs
are replayed through the oracle (CPU suite) and through the CUDA path via the C ABI (GPU suite)."""
import glob
import os

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200 import abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WINDOWS = ["c1_single_scan", "c2_small", "tracking2_segments", "init_segments"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    fields = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    hb = abi.HostBatch(int(z["n_windows"]), int(z["n_frames"]), int(z["ground_multiplicity"]), int(z["prior_frame"]), **fields)
    return z, hb, L.corridor_params(max_iters=int(z["max_iters"]))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def test_golden_files_present():
    assert sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLD, "*.npz"))) == sorted([w + ".npz" for w in WINDOWS] + ["factors.npz", "lines_scans.npz", "pose_graph.npz", "ref_text.npz"])


@pytest.mark.parametrize("name", WINDOWS)
def test_oracle_reproduces_golden(oracle, name):
    z, hb, P = load(name)
    for mode in (0, 1):
        H, g, c = oracle.linearize(P, hb, mode=mode)
        assert rel(H, z[f"H{mode}"]) < 1e-10 and rel(g, z[f"g{mode}"]) < 1e-10 and rel(c, z[f"cost{mode}"]) < 1e-12
    st, summ = oracle.solve(P, hb)
    assert np.array_equal(summ["iterations"], z["summary_iterations"])
    assert np.abs(st - z["solved_states"]).max() < 1e-8
    if "raw_imu_samples" in z.files:
        blobs = oracle.imu_preintegrate(P, z["raw_imu_offset"], z["raw_imu_samples"], z["raw_bias0"])
        assert rel(blobs, hb["imu"].reshape(blobs.shape)) < 1e-9


def test_oracle_factors_reproduce_golden(oracle):
    z = np.load(os.path.join(GOLD, "factors.npz"))
    P = L.corridor_params()
    r, J = oracle.eval_imu_factor(P, z["imu_blob"], z["state_i"], z["state_j"])
    assert rel(r, z["imu_res"]) < 1e-11 and rel(J, z["imu_jac"]) < 1e-11
    r, J = oracle.eval_wheel_factor(P, z["wheel_blob"], z["state_i"][:6], z["state_j"][:6])
    assert rel(r, z["wheel_res"]) < 1e-11 and rel(J, z["wheel_jac"]) < 1e-10
    l = z["laser_lines"]
    r, J = oracle.eval_laser_factor(P, l[0], l[1], l[2], l[3], z["state_i"][:6], z["laser_pose_j"])
    assert rel(r, z["laser_res"]) < 1e-12 and rel(J, z["laser_jac"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("name", WINDOWS)
def test_gpu_reproduces_golden(name):
    from lvio2d_b200.solver import Context

    z, hb, P = load(name)
    with Context(P) as c:
        c.set_windows(hb)
        for mode in (0, 1):
            H, g, cost = c.linearize(mode)
            assert rel(H, z[f"H{mode}"]) < 1e-9 and rel(g, z[f"g{mode}"]) < 1e-9 and rel(cost, z[f"cost{mode}"]) < 1e-11
        summ = c.solve()
        st = c.get_states()
        assert np.array_equal(summ["iterations"], z["summary_iterations"])
        assert np.array_equal(summ["termination"], z["summary_termination"])
        assert np.abs(st - z["solved_states"]).max() < 1e-8
        assert rel(summ["final_cost"], z["summary_final_cost"]) < 1e-8
        X0, J, r = c.marginalize()   # at the solved states: compare with the oracle's marginalisation of the same states
    # marginalisation fixture was taken at the initial states
    with Context(P) as c:
        c.set_windows(hb)
        X0, J, r = c.marginalize()
        JTJ = np.einsum("bki,bkj->bij", J, J)
        gJTJ = np.einsum("bki,bkj->bij", z["marg_J"], z["marg_J"])
        assert rel(JTJ, gJTJ) < 1e-7
        assert np.array_equal(X0, z["marg_X0"])
        if "raw_imu_samples" in z.files:
            blobs = c.imu_preintegrate(z["raw_imu_offset"], z["raw_imu_samples"], z["raw_bias0"])
            assert rel(blobs[:, :240], hb["imu"].reshape(blobs.shape)[:, :240]) < 1e-10
            assert rel(blobs[:, 240:465], hb["imu"].reshape(blobs.shape)[:, 240:465]) < 1e-7


@pytest.mark.gpu
def test_gpu_factors_reproduce_golden():
    from lvio2d_b200.solver import Context

    z = np.load(os.path.join(GOLD, "factors.npz"))
    with Context(L.corridor_params()) as c:
        r, J = c.eval_imu_factor(z["imu_blob"], z["state_i"], z["state_j"])
        assert rel(r, z["imu_res"]) < 1e-9 and rel(J, z["imu_jac"]) < 1e-9
        r, J = c.eval_wheel_factor(z["wheel_blob"], z["state_i"][:6], z["state_j"][:6])
        assert rel(r, z["wheel_res"]) < 1e-9 and rel(J, z["wheel_jac"]) < 1e-8
        r, J = c.eval_ground_factors(z["state_j"][:6])
        assert rel(r, z["ground_res"]) < 1e-9 and rel(J, z["ground_jac"]) < 1e-9
        l = z["laser_lines"]
        r, J = c.eval_laser_factor(l[0], l[1], l[2], l[3], z["state_i"][:6], z["laser_pose_j"])
        assert rel(r, z["laser_res"]) < 1e-10 and rel(J, z["laser_jac"]) < 1e-9


# ---- back-end pose graph (tests/golden/pose_graph.npz, scripts/make_golden_pose_graph.py)
POSE_GRAPHS = ["ring24", "loops40", "loops40_ground_q_12it"]


def load_pose_graph(name):
    z = np.load(os.path.join(GOLD, "pose_graph.npz"))
    g = {k.split("__", 1)[1]: z[k] for k in z.files if k.startswith(name + "__")}
    return g, z["sqrt_info"], L.corridor_params(max_iters=int(g["max_iters"]))


@pytest.mark.parametrize("name", POSE_GRAPHS)
def test_oracle_reproduces_golden_pose_graph(oracle, name):
    g, Jn, P = load_pose_graph(name)
    x, s = oracle.pose_graph_solve(P, g["init"], g["edges"], g["tfs"], g["weights"], Jn, ground_p=True, ground_q=bool(g["ground_q"]))
    assert np.array_equal(s["iterations"], g["iterations"]) and np.array_equal(s["termination"], g["termination"])
    assert rel(s["final_cost"], g["final_cost"]) < 1e-9 and np.abs(x - g["solved"]).max() < 1e-9
    e = len(g["edges"]) - 1
    r, J = oracle.eval_edge_factor(g["tfs"][e], g["weights"][e], Jn, g["init"][g["edges"][e][0]], g["init"][g["edges"][e][1]])
    assert rel(r, g["last_edge_res"]) < 1e-12 and rel(J, g["last_edge_jac"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", POSE_GRAPHS)
def test_cuda_reproduces_golden_pose_graph(name):
    from lvio2d_b200.solver import Context

    g, Jn, P = load_pose_graph(name)
    with Context(P) as ctx:
        x, s = ctx.pose_graph_solve(g["init"], g["edges"], g["tfs"], g["weights"], Jn, True, bool(g["ground_q"]))
        e = len(g["edges"]) - 1
        r, J = ctx.eval_edge_factor(g["tfs"][e], g["weights"][e], Jn, g["init"][g["edges"][e][0]], g["init"][g["edges"][e][1]])
    assert np.array_equal(s["termination"], g["termination"]) and abs(int(s["iterations"][0]) - int(g["iterations"][0])) <= 1
    assert rel(s["final_cost"], g["final_cost"]) < 1e-6 and np.abs(x - g["solved"]).max() < 1e-6
    assert rel(r, g["last_edge_res"]) < 1e-11 and rel(J, g["last_edge_jac"]) < 1e-9
