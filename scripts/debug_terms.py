import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import lvio2d_b200 as L
import oracle_lib as O
from lvio2d_b200.solver import Context
P = L.corridor_params(max_iters=10)
sb = L.synth.make_batch(2, 42, n_frames=5, beams=300, fov_deg=270.0)
hb = O.preintegrate_batch(P, sb)
variants = {
 "all": hb,
 "no_ground": hb.replace(ground_multiplicity=0),
 "no_laser": hb.replace(point_offset=None, points=None, point_line=None, line_offset=None, lines=None, ref_frame=None, ref_pose=None),
 "no_prior": hb.replace(prior_frame=-1, prior_X0=None, prior_J=None),
 "no_imu": hb.replace(imu=None),
 "no_wheel": hb.replace(wheel=None),
}
with Context(P) as c:
    for mode in (0, 1):
        for name, b in variants.items():
            c.set_windows(b)
            H, g, cost = c.linearize(mode)
            oH, og, oc = O.linearize(P, b, mode=mode)
            print(mode, name, "cost", cost, oc, "dcost", cost - oc, "g", np.abs(g - og).max() / np.abs(og).max(), "H", np.abs(H - oH).max() / np.abs(oH).max())
