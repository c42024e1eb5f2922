"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import lvio2d_b200 as L
import oracle_lib as O
from lvio2d_b200.solver import Context
for assoc, huber in ((0, 0.0), (1, 1.5)):
    P = L.corridor_params(max_iters=4); P.assoc_mode = assoc; P.huber_delta = huber
    for mk in (lambda: L.synth.make_batch(2, 42, n_frames=4, beams=90), lambda: L.synth.config_init(1, n_frames=4), lambda: L.synth.config_tracking2(1), lambda: L.synth.config_c1()):
        sb = mk()
        with Context(P) as c:
            hb = c.preintegrate_batch(sb)
            c.set_windows(hb)
            c.linearize(0); c.linearize(1)
            s = c.solve(); x = c.get_states()
            c.marginalize()
            c.set_point_shard(0, 2); c.solve_begin(); c.eval_laser(); c.lm_step(want_active=True); c.set_point_shard(0, 1)
        print("ok", assoc, huber, hb.n_frames, s["final_cost"])
