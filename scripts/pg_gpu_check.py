"""One-shot GPU check of the pose-graph device path against the oracle (what tests/test_zz_gpu_pose_graph.py asserts,
without pytest start-up), plus wall-clock timings of larger graphs.  Writes gpurun_out/pg_check.txt.
    gpurun --timeout 120 -- 'python scripts/pg_gpu_check.py'"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import lvio2d_b200 as L  # noqa: E402
import oracle_lib as O  # noqa: E402
from lvio2d_b200.solver import Context  # noqa: E402
from test_oracle_pose_graph import T_of, edge_noise_J  # noqa: E402
from test_pose_graph_host import graph_with_loops  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "pg_check.txt"), "w")


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    out.write(line + "\n")
    out.flush()


Jn = edge_noise_J()
g = np.random.default_rng(3)
with Context(L.corridor_params()) as ctx:
    worst_r = worst_J = 0.0
    for case in range(20):
        pi, pj = np.r_[g.uniform(-5, 5, 3), g.normal(0, 0.5, 3)], np.r_[g.uniform(-5, 5, 3), g.normal(0, 0.5, 3)]
        noise = np.r_[g.normal(0, 0.05, 3), g.normal(0, 0.02, 3)] * (1e-4 if case % 2 else 1.0)
        tf12 = np.ascontiguousarray((np.linalg.inv(T_of(pi)) @ T_of(pj) @ T_of(noise))[:3, :])
        res, jac = ctx.eval_edge_factor(tf12, 2.0, Jn, pi, pj)
        wr, wJ = O.eval_edge_factor(tf12, 2.0, Jn, pi, pj)
        worst_r = max(worst_r, np.abs(res - wr).max() / max(1.0, np.abs(wr).max()))
        worst_J = max(worst_J, np.abs(jac - wJ).max() / np.abs(wJ).max())
    say("edge_factor worst rel err: res", worst_r, "jac", worst_J, "PASS" if worst_r < 1e-12 and worst_J < 1e-9 else "FAIL")

cases = [(24, [], False, 50), (24, [], True, 12), (40, [(30, 4), (12, 25), (39, 20)], False, 50), (40, [(30, 4), (12, 25), (39, 20)], True, 12),
         (200, [(190, 3), (100, 20), (150, 40), (199, 80), (60, 160)], False, 50)]
for K, loops, gq, iters in cases:
    P = L.corridor_params(max_iters=iters)
    truth, init, edges, tfs, ws = graph_with_loops(K, loops, seed=4 + K)
    want, s1 = O.pose_graph_solve(P, init, edges, tfs, ws, Jn, ground_p=True, ground_q=gq)
    with Context(P) as ctx:
        ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, gq)     # warm-up (allocations)
        t0 = time.perf_counter()
        got, s2 = ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, gq)
        dt = time.perf_counter() - t0
    err = np.abs(got - want).max()
    say(f"K={K} L={len(loops) + 1} ground_q={gq} iters oracle/gpu {s1['iterations'][0]}/{s2['iterations'][0]} term {s1['termination'][0]}/{s2['termination'][0]}"
        f" cost {s1['final_cost'][0]:.9g}/{s2['final_cost'][0]:.9g} max|dpose| {err:.3g} wall {dt * 1e3:.2f} ms", "PASS" if err < 1e-6 else "FAIL")
# timings only (the dense oracle is too slow here)
for K, nl in ((1000, 8), (4000, 16)):
    loops = [(K - 10 - 7 * i, 5 + 11 * i) for i in range(nl)]
    truth, init, edges, tfs, ws = graph_with_loops(K, loops, seed=K)
    with Context(L.corridor_params(max_iters=50)) as ctx:
        ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
        t0 = time.perf_counter()
        got, s2 = ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
        dt = time.perf_counter() - t0
    say(f"K={K} L={nl + 1}: {s2['iterations'][0]} LM iterations, cost {s2['initial_cost'][0]:.6g} -> {s2['final_cost'][0]:.6g}, wall {dt * 1e3:.1f} ms"
        f" ({dt * 1e3 / max(1, s2['iterations'][0]):.2f} ms / iteration), max |pose - truth| {np.abs(got[:, :3] - truth[:, :3]).max():.3g} m")
# the opt-in partitioned solve (pose_graph_segments.cuh): parity against the plain path and its wall clock
for K, nl, P, stage in ((200, 5, 16, 0), (1000, 8, 64, 0), (4000, 16, 64, 0), (4000, 16, 128, 0), (1000, 8, 32, 1), (4000, 16, 64, 1), (4000, 16, 32, 1),
                        (1000, 8, "auto", 1), (4000, 16, "auto", 1)):
    loops = [(K - 10 - 7 * i, 5 + 11 * i) for i in range(nl)]
    truth, init, edges, tfs, ws = graph_with_loops(K, loops, seed=K)
    os.environ.pop("LVIO2D_PG_SEGMENTS", None)
    with Context(L.corridor_params(max_iters=50)) as ctx:
        plain, s0 = ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
    os.environ["LVIO2D_PG_SEGMENTS"] = str(P)
    os.environ["LVIO2D_PG_STAGE"] = str(stage)
    with Context(L.corridor_params(max_iters=50)) as ctx:
        ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
        t0 = time.perf_counter()
        got, s2 = ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
        dt = time.perf_counter() - t0
    os.environ.pop("LVIO2D_PG_SEGMENTS", None)
    os.environ.pop("LVIO2D_PG_STAGE", None)
    err = np.abs(got - plain).max()
    say(f"segments={P} stage={stage} K={K} L={nl + 1}: iterations plain/partitioned {s0['iterations'][0]}/{s2['iterations'][0]}, max|dpose| {err:.3g}, wall {dt * 1e3:.1f} ms"
        f" ({dt * 1e3 / max(1, s2['iterations'][0]):.2f} ms / iteration)", "PASS" if err < 1e-7 else "FAIL")
out.close()
