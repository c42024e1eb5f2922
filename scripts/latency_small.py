"""Latency of ONE solve for the reference's real per-frame problem (2 frames, matched segment pairs) and for one C2 window."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import lvio2d_b200 as L
import oracle_lib as O
from lvio2d_b200.solver import Context
for name, mk, iters in (("tracking2 (reference per-frame problem)", lambda: L.synth.config_tracking2(1), 50), ("init n=10 segments", lambda: L.synth.config_init(1), 50), ("C2 1 window", lambda: L.synth.config_c2(1), 10)):
    P = L.corridor_params(max_iters=iters)
    sb = mk()
    hb = O.preintegrate_batch(P, sb)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps): st, summ = O.solve(P, hb)
    cpu_ms = (time.perf_counter() - t0) / reps * 1e3
    with Context(P) as c:
        for _ in range(3):
            c.set_windows(hb); c.solve(False); c.get_states()
        t0 = time.perf_counter(); reps = 20
        for _ in range(reps):
            c.set_windows(hb); c.solve(False); x = c.get_states()
        gpu_ms = (time.perf_counter() - t0) / reps * 1e3
        s2 = c.get_summaries()
    print(f"{name}: points {hb.n_points}, iterations cpu {int(summ['iterations'][0])} gpu {int(s2['iterations'][0])}; oracle CPU {cpu_ms:.2f} ms/solve, GPU end-to-end {gpu_ms:.2f} ms/solve")
