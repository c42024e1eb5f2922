"""Golden vectors written by the REFERENCE'S OWN source text (oracle/_ref/libref.so: src/factor/*.h, solver.cpp,
laser_manager.cpp, common.cpp of /root/reference compiled unmodified against the stub include tree, oracle/Makefile).
Run here, where /root/reference exists; the fixture tests/golden/ref_text.npz travels and pins the oracle (CPU tests)
and the CUDA path (GPU tests) to the reference text even where libref.so cannot be built.
    python scripts/make_golden_ref.py"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

import lvio2d_b200 as L  # noqa: E402
import oracle_lib as O  # noqa: E402
import ref_lib  # noqa: E402
from lvio2d_b200 import replay  # noqa: E402
from lvio2d_b200.params import params_T  # noqa: E402

g = np.random.default_rng(20261018)
P = L.corridor_params()
T_io = params_T(P, "T_imu_to_wheel")


def corridor_pose():
    Rwb = Rotation.from_euler("zyx", [g.uniform(-3, 3), g.normal(0, 2e-3), g.normal(0, 2e-3)]).as_matrix()
    return np.concatenate([g.uniform(-5, 5, 3), Rotation.from_matrix(Rwb @ T_io[:, :3].T).as_rotvec()])


def state(pose):
    return np.concatenate([pose, g.normal(0, 0.5, 3), g.normal(0, 0.02, 3), g.normal(0, 0.002, 3)])


out = {}
N = 12
# laser factor
pi_ = np.array([corridor_pose() for _ in range(N)])
pj_ = pi_.copy()
pj_[:, 0:3] += g.normal(0, 0.3, (N, 3))
for k in range(N):
    pj_[k, 3:6] = (Rotation.from_rotvec(pi_[k, 3:6]) * Rotation.from_rotvec(g.normal(0, 0.05, 3))).as_rotvec()
ends = np.concatenate([g.uniform(-6, 6, (N, 4, 2)), np.zeros((N, 4, 1))], axis=2)
lr, lJ = zip(*[ref_lib.eval_laser_factor(*ends[k], pi_[k], pj_[k]) for k in range(N)])
out.update(laser_ends=ends, laser_pose_i=pi_, laser_pose_j=pj_, laser_res=np.array(lr), laser_jac=np.array(lJ))
# IMU: preintegration + factor
samples = np.zeros((N, 20, 7))
samples[:, :, 0] = g.uniform(0.002, 0.006, (N, 20))
samples[:, :, 1:4] = g.normal(0, 1.0, (N, 20, 3)) + np.array([0.2, 9.7, 0.1])
samples[:, :, 4:7] = g.normal(0, 0.3, (N, 20, 3))
bias = np.concatenate([g.normal(0, 0.02, (N, 3)), g.normal(0, 0.002, (N, 3))], axis=1)
blobs = np.array([ref_lib.imu_preintegrate([0, 20], samples[k], bias[k][None])[0] for k in range(N)])
si = np.array([state(corridor_pose()) for _ in range(N)])
si[:, 9:15] = bias
sj = si.copy()
for k in range(N):
    sj[k, 0:3] += si[k, 6:9] * blobs[k, 465] + g.normal(0, 0.02, 3)
    sj[k, 3:6] = (Rotation.from_rotvec(si[k, 3:6]) * Rotation.from_rotvec(blobs[k, 6:9] + g.normal(0, 0.01, 3))).as_rotvec()
    sj[k, 6:9] += g.normal(0, 0.1, 3)
    sj[k, 9:15] += g.normal(0, 1e-3, 6)
ir, iJ = zip(*[ref_lib.eval_imu_factor(blobs[k], si[k], sj[k]) for k in range(N)])
out.update(imu_samples=samples, imu_bias=bias, imu_blob=blobs, imu_si=si, imu_sj=sj, imu_res=np.array(ir), imu_jac=np.array(iJ))
# wheel: preintegration + factor (regular branch)
steps = np.zeros((N, 5, 7))
steps[:, :, 0] = 0.02
steps[:, :, 1] = g.uniform(0.2, 1.0, (N, 1))
steps[:, :, 6] = g.normal(0, 0.5, (N, 1))
wblobs = np.array([ref_lib.wheel_preintegrate([0, 5], steps[k])[0] for k in range(N)])
wpi = np.array([corridor_pose() for _ in range(N)])
wpj = wpi.copy()
Tio4 = np.eye(4)
Tio4[:3, :4] = T_io
for k in range(N):
    Ri = Rotation.from_rotvec(wpi[k, 3:6]).as_matrix()
    dT = np.eye(4)
    dT[:3, :4] = wblobs[k, 0:12].reshape(3, 4)
    Two = np.eye(4)
    Two[:3, :3], Two[:3, 3] = Ri @ T_io[:, :3], Ri @ T_io[:, 3] + wpi[k, 0:3]
    Tj = Two @ dT @ np.linalg.inv(Tio4)
    wpj[k] = np.concatenate([Tj[:3, 3] + g.normal(0, 1e-3, 3), Rotation.from_matrix(Tj[:3, :3]).as_rotvec()])
wr, wJ = zip(*[ref_lib.eval_wheel_factor(wblobs[k], wpi[k], wpj[k]) for k in range(N)])
out.update(wheel_steps=steps, wheel_blob=wblobs, wheel_pose_i=wpi, wheel_pose_j=wpj, wheel_res=np.array(wr), wheel_jac=np.array(wJ))
# ground
gp = np.array([corridor_pose() for _ in range(N)])
for k in range(N):
    gp[k, 3:6] = (Rotation.from_rotvec(gp[k, 3:6]) * Rotation.from_rotvec(g.normal(0, 0.05, 3))).as_rotvec()
gr, gJ = zip(*[ref_lib.eval_ground_factors(gp[k]) for k in range(N)])
out.update(ground_pose=gp, ground_res=np.array(gr), ground_jac=np.array(gJ))
# solver::solve (fast_mode: 10 iterations) and marginalization on a synthetic 6-frame window
Pf = L.corridor_params(fast_mode=True)
sb = replay.make_sequence(5, n_frames=8, params=Pf)
hb = O.preintegrate_batch(Pf, sb)
frames = replay.frames_of(sb, hb["imu"], hb["wheel"])
for n in (2, 6):
    a = copy.deepcopy(frames[:n])
    ref = ref_lib.RefSolver(fast_mode=True)
    ref.solve(a)
    out[f"solve{n}_states"] = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in a])
    out[f"solve{n}_summary"] = np.array([int(ref.last_summary["iterations"][0]), int(ref.last_summary["termination"][0]),
                                         float(ref.last_summary["initial_cost"][0]), float(ref.last_summary["final_cost"][0])])
a = copy.deepcopy(frames[:3])
refd = ref_lib.RefSolver(fast_mode=False)
refd.marginalization(a)
X0, J, r = refd.prior
out.update(marg_X0=X0, marg_JTJ=J.T @ J)
# laser front-end: spawn_scan and do_match
lp = L.corridor_line_params()
off, pts = L.synth.make_scan_batch(3, 11, beams=721, range_sigma=0.004)
for k in range(3):
    p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
    sc = ref_lib.RefScan.from_points(p3)
    out[f"scan{k}_lines"] = sc.lines()[0]
s1 = ref_lib.RefScan.from_points(np.c_[pts[off[0]:off[1]], np.zeros(off[1] - off[0])])
T_il = np.array(list(P.T_imu_to_laser)).reshape(3, 4)
base = Rotation.from_matrix(T_il[:, :3].T)
pose1 = np.r_[0.5, -0.3, 0.0, base.as_rotvec()]
Rl = Rotation.from_euler("z", 0.015).as_matrix()
dxy = np.array([0.03, -0.02])
R1 = base.as_matrix()
Rwl1, twl1 = R1 @ T_il[:, :3], R1 @ T_il[:, 3] + pose1[0:3]
Rwl2, twl2 = Rwl1 @ Rl, twl1 + Rwl1 @ np.r_[dxy, 0.0]
R2 = Rwl2 @ T_il[:, :3].T
pose2 = np.r_[twl2 - R2 @ T_il[:, 3], Rotation.from_matrix(R2).as_rotvec()]
q3 = (Rl.T @ (np.c_[pts[off[0]:off[1]], np.zeros(off[1] - off[0])] - np.r_[dxy, 0.0]).T).T
s2 = ref_lib.RefScan.from_points(q3)
out.update(match_points2=q3, match_pose1=pose1, match_pose2=pose2, match_lines2=s2.lines()[0],
           match_pairs_kk0=ref_lib.do_match(s1, s2, pose1, pose2, 0), match_pairs_kk1=ref_lib.do_match(s1, s2, pose1, pose2, 1))
path = os.path.join(ROOT, "tests", "golden", "ref_text.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
