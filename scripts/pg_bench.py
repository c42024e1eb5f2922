"""Back-end pose-graph row timed through the C ABI with host buffers (what keyframe_manager::solve
would call): wall clock of lvio2d_pose_graph_solve per LM iteration on synthetic ring graphs, next to the CPU oracle
(dense normal equations, one core) on the size it can finish in seconds.  Prints ONE JSON object; run by bench.py in a
subprocess (`next_rows.pose_graph`) so that nothing here can take the headline line down.
    python scripts/pg_bench.py [--no-cpu]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import lvio2d_b200 as L  # noqa: E402
from lvio2d_b200.backend import edge_noise_J  # noqa: E402
from lvio2d_b200.solver import Context  # noqa: E402


def ring_graph(K, loops, seed):
    """Key frames on a 3 m circle in the ground plane, noisy sequential edges, the ring closure + extra loop edges."""
    from scipy.spatial.transform import Rotation

    from lvio2d_b200.params import params_T

    g = np.random.default_rng(seed)
    T_io = params_T(L.corridor_params(), "T_imu_to_wheel")

    def T_of(p):
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = Rotation.from_rotvec(p[3:6]).as_matrix(), p[0:3]
        return T

    truth = np.zeros((K, 6))
    for k in range(K):
        a = 2 * np.pi * k / K
        R_wb = Rotation.from_euler("z", a + np.pi / 2).as_matrix()
        R_wi, p_wi = L.synth._imu_pose_from_base(R_wb, np.array([3 * np.cos(a), 3 * np.sin(a), 0.0]), T_io)
        truth[k, 0:3], truth[k, 3:6] = p_wi, Rotation.from_matrix(R_wi).as_rotvec()
    Ts = [T_of(x) for x in truth]
    edges, tfs, ws = [], [], []
    for k in range(K - 1):
        noise = T_of(np.r_[g.normal(0, 0.02, 3), g.normal(0, 0.005, 3)])
        edges.append((k, k + 1)); tfs.append((np.linalg.inv(Ts[k]) @ Ts[k + 1] @ noise)[:3, :]); ws.append(1.0)
    for (i, j) in [(K - 1, 0)] + list(loops):
        noise = T_of(np.r_[g.normal(0, 0.005, 3), g.normal(0, 0.002, 3)])
        edges.append((i, j)); tfs.append((np.linalg.inv(Ts[i]) @ Ts[j] @ noise)[:3, :]); ws.append(10.0)
    init = truth.copy()
    T = Ts[0]
    for k in range(K - 1):
        T = T @ np.vstack([tfs[k], [0, 0, 0, 1]])
        init[k + 1] = np.r_[T[:3, 3], Rotation.from_matrix(T[:3, :3]).as_rotvec()]
    return init, np.array(edges, np.int32), np.array(tfs), np.array(ws)


def cpu_port():
    """The same algorithm (block-tridiagonal + loop-edge capacitance, same minimiser loop) as scalar C++ on one host core:
    tests/native/pose_graph_host.cpp — the test suite's thread-by-thread host run of the kernel bodies — built with -O3.
    A fairer CPU number than the dense oracle; the reference's own SPARSE_SCHUR cannot be built here."""
    import ctypes as C
    import subprocess
    import tempfile

    from test_device_math_host import Consts

    so = os.path.join(tempfile.mkdtemp(prefix="pgh_"), "libpgh.so")
    subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++",
                           os.path.join(ROOT, "tests", "native", "pose_graph_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    P = L.corridor_params()
    c = Consts()
    c.T_il[:] = list(P.T_imu_to_laser)
    c.T_io[:] = list(P.T_imu_to_wheel)
    c.g = P.g
    c.laser_sqrt_info, c.ground_p_sqrt_info, c.ground_q_sqrt_info = 1.0 / P.line_to_line_sigma, 1.0 / P.manifold_p_sigma, 1.0 / P.manifold_q_sigma

    def solve(init, edges, tfs, ws, Jn):
        from lvio2d_b200 import abi

        dp = C.POINTER(C.c_double)
        x = np.array(init, dtype=np.float64).copy()
        ei = np.ascontiguousarray(edges, dtype=np.int32)
        et, ew, J = np.ascontiguousarray(tfs, dtype=np.float64), np.ascontiguousarray(ws, dtype=np.float64), np.ascontiguousarray(Jn, dtype=np.float64)
        opt = np.array([50, 1e-6, 1e-10, 1e-8, 1e4])
        summ = np.zeros(1, dtype=abi.SUMMARY_DTYPE)
        t0 = time.perf_counter()
        rc = lib.pgh_solve(C.byref(c), opt.ctypes.data_as(dp), len(x), x.ctypes.data_as(dp), len(ei), ei.ctypes.data_as(C.POINTER(C.c_int32)),
                           et.ctypes.data_as(dp), ew.ctypes.data_as(dp), J.ctypes.data_as(dp), 1, 0, summ.ctypes.data_as(C.c_void_p), None, 0, 0)
        dt = time.perf_counter() - t0
        assert rc == 0
        return x, summ, dt

    return solve


def main():
    cpu = "--no-cpu" not in sys.argv
    port = cpu_port() if cpu else None
    Jn = edge_noise_J((0.1,) * 3, (0.01,) * 3)
    out = {"unit": "ms per LM iteration (wall clock of lvio2d_pose_graph_solve, host buffers, ground_p on, ground_q off)", "graphs": []}
    for K, nl in ((200, 5), (1000, 8), (4000, 16)):
        loops = [(K - 10 - 7 * i, 5 + 11 * i) for i in range(nl)]
        init, edges, tfs, ws = ring_graph(K, loops, seed=K)
        P = L.corridor_params(max_iters=50)
        with Context(P) as ctx:
            ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)            # allocations
            t0 = time.perf_counter()
            got, s = ctx.pose_graph_solve(init, edges, tfs, ws, Jn, True, False)
            dt = time.perf_counter() - t0
        row = {"key_frames": K, "loop_edges": nl + 1, "iterations": int(s["iterations"][0]), "wall_ms": round(dt * 1e3, 3),
               "ms_per_iteration": round(dt * 1e3 / max(1, int(s["iterations"][0])), 4), "final_cost": float(s["final_cost"][0])}
        if cpu and K <= 200:
            import oracle_lib as O

            t0 = time.perf_counter()
            want, so = O.pose_graph_solve(P, init, edges, tfs, ws, Jn, ground_p=True, ground_q=False)
            dc = time.perf_counter() - t0
            row["cpu_oracle_dense_1core"] = {"wall_ms": round(dc * 1e3, 3), "iterations": int(so["iterations"][0]),
                                              "max_abs_pose_diff": float(np.abs(got - want).max())}
        if port is not None:
            xp, sp, dp_ = port(init, edges, tfs, ws, Jn)
            row["cpu_port_same_algorithm_1core"] = {"wall_ms": round(dp_ * 1e3, 3), "iterations": int(sp["iterations"][0]),
                                                    "ms_per_iteration": round(dp_ * 1e3 / max(1, int(sp["iterations"][0])), 4),
                                                    "max_abs_pose_diff": float(np.abs(got - xp).max())}
        out["graphs"].append(row)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
