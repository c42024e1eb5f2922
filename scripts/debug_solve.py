import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import lvio2d_b200 as L
import oracle_lib as O
from lvio2d_b200.solver import Context
for case, mk in [("tracking2", lambda: L.synth.config_tracking2(2)), ("init", lambda: L.synth.config_init(2, n_frames=6)), ("init10", lambda: L.synth.config_init(1, n_frames=10))]:
    for iters in (5, 10, 20, 50):
        P = L.corridor_params(max_iters=iters)
        sb = mk()
        hb = O.preintegrate_batch(P, sb)
        with Context(P) as c:
            c.set_windows(hb); summ = c.solve(); got = c.get_states()
        want, osumm = O.solve(P, hb)
        d = np.abs(got - want)
        n = hb.n_frames
        g3 = got.reshape(-1, n, 15); w3 = want.reshape(-1, n, 15)
        rel = np.abs((g3[:, :, 0:3] - g3[:, :1, 0:3]) - (w3[:, :, 0:3] - w3[:, :1, 0:3])).max()
        print(case, iters, "succ", summ["num_successful_steps"], osumm["num_successful_steps"], "cost", summ["final_cost"], osumm["final_cost"],
              "dp %.2e dq %.2e dv %.2e dbs %.2e relp %.2e" % (d[:, 0:3].max(), d[:, 3:6].max(), d[:, 6:9].max(), d[:, 9:].max(), rel),
              "err_truth %.2e" % np.abs(got - sb.truth)[:, 0:3].max())
