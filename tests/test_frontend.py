"""The host-side mirror of the reference's laser front-end interface (lvio2d_b200.frontend: Laser, LaserManager.spawn_scan,
LaserManager.do_match) in the call order of trajectory::add_sensor_data (reference src/trajectory/trajectory.cpp:136-260):
LaserScan message -> sensor::laser -> correct() -> spawn_scan -> do_match -> frame_info::add_laser_match -> solver::solve.
CPU: the flow on the oracle backend.  GPU: the same flow on the CUDA library against it."""
import math

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200.frontend import Laser, LaserManager
from lvio2d_b200.solver import FrameInfo, Solver

BEAMS, FOV = 1081, 270.0


def _messages(seed=3):
    """Two LaserScan messages of one synthetic world seen from two nearby poses, with the IMU poses and blobs."""
    sb = L.synth.make_batch(1, seed, n_frames=2, beams=BEAMS, fov_deg=FOV, n_segments=12, frame_dt=0.3)
    assert np.all(np.diff(sb.point_offset) == BEAMS)           # every beam valid: the points are the beams in order
    a0 = np.float32(math.radians(-FOV / 2.0))
    da = np.float32(math.radians(FOV) / (BEAMS - 1))
    msgs = []
    for k in range(2):
        pts = sb.points[sb.point_offset[k]:sb.point_offset[k + 1]]
        msgs.append(dict(ranges=np.linalg.norm(pts, axis=1).astype(np.float32), angle_min=a0, angle_increment=da,
                         time_increment=np.float32(0.025 / BEAMS), stamp=100.0 + 0.3 * k))
    return sb, msgs


def _flow(backend, solver, sb, msgs, hb):
    lm = LaserManager(backend, L.corridor_line_params(), max_lines=160)
    scans = []
    for k, m in enumerate(msgs):
        laser = Laser(backend, **m)
        laser.correct(np.array([0.3, 0.02, 0.0]) * k, np.array([0.0, 0.0, 0.1]) * k)     # the second scan is taken on the move
        scans.append(lm.spawn_scan(laser))
    match = lm.do_match(scans[0], scans[1], sb.truth[0, 0:3], sb.truth[0, 3:6], sb.states[1, 0:3], sb.states[1, 3:6])
    f0 = FrameInfo(0.0, *np.split(sb.truth[0], [3, 6, 9]))
    s1 = sb.states[1]
    f1 = FrameInfo(0.3, s1[0:3], s1[3:6], s1[6:9], s1[9:15], hb["imu"].reshape(-1, 466)[0], hb["wheel"].reshape(-1, 15)[0])
    f1.add_laser_match(match)
    solver.solve([f0, f1])
    return scans, match, np.r_[f1.p, f1.q]


def test_front_end_flow_on_the_oracle_backend(oracle):
    P = L.corridor_params(max_iters=10)
    sb, msgs = _messages()
    hb = oracle.preintegrate_batch(P, sb)
    be = oracle.OracleContext(P)
    scans, match, pose = _flow(be, Solver(P, fast_mode=True, ctx=be), sb, msgs, hb)
    assert len(scans[0].lines) > 10 and len(scans[1].lines) > 10
    assert len(match.lines1) == len(match.lines2) >= 5
    assert all(l.index2 - l.index1 >= 2 for s in scans for l in s.lines)
    for l1, l2 in zip(match.lines1, match.lines2):              # matched segments are nearly parallel (< 10 degrees)
        assert l1 in scans[0].lines and l2 in scans[1].lines
    # the solve moved the new frame towards the truth
    assert np.abs(pose - sb.truth[1, 0:6]).max() < np.abs(sb.states[1, 0:6] - sb.truth[1, 0:6]).max()


@pytest.mark.gpu
def test_front_end_flow_matches_oracle(oracle):
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=10)
    sb, msgs = _messages()
    hb = oracle.preintegrate_batch(P, sb)
    be = oracle.OracleContext(P)
    oscans, omatch, opose = _flow(be, Solver(P, fast_mode=True, ctx=be), sb, msgs, hb)
    with Context(P) as c:
        scans, match, pose = _flow(c, Solver(P, fast_mode=True, ctx=c), sb, msgs, hb)
    for s, o in zip(scans, oscans):
        assert [(l.index1, l.index2) for l in s.lines] == [(l.index1, l.index2) for l in o.lines]
        assert np.abs(s.points - o.points).max() < 1e-12
    assert [(scans[0].lines.index(a), scans[1].lines.index(b)) for a, b in zip(match.lines1, match.lines2)] == \
           [(oscans[0].lines.index(a), oscans[1].lines.index(b)) for a, b in zip(omatch.lines1, omatch.lines2)]
    assert np.abs(pose - opose).max() < 1e-6
