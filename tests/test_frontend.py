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


def _scan_sequence(oracle, n=8, seed=11):
    """n scans of one world along a short trajectory (synthetic window frames), as Scan objects + true IMU poses."""
    P = L.corridor_params()
    sb = L.synth.make_batch(1, seed, n_frames=n, beams=BEAMS, fov_deg=FOV, n_segments=12, frame_dt=0.3)
    be = oracle.OracleContext(P)
    lm = LaserManager(be, L.corridor_line_params(), max_lines=160, params=P, ref_n_accumulation=6)
    a0, da = np.float32(math.radians(-FOV / 2.0)), np.float32(math.radians(FOV) / (BEAMS - 1))
    scans = []
    for k in range(n):
        pts = sb.points[sb.point_offset[k]:sb.point_offset[k + 1]]
        laser = Laser(be, np.linalg.norm(pts, axis=1).astype(np.float32), a0, da, np.float32(0.0), 100.0 + 0.3 * k)
        scans.append(lm.spawn_scan(laser))
    return P, sb, lm, scans


def test_add_scan_bookkeeping_and_match_with_ref(oracle):
    """laser_manager::add_scan / match_with_ref (laser_manager.cpp:424-547) as host logic: first scan initialises the
    reference sub-map, the motion filter skips a scan that did not move, lines arrive in the sub-map's laser frame, the
    spawning sub-map appears at n/2 and replaces the reference at n; matching a new scan against the sub-map pairs
    segments that coincide once both are taken to the world."""
    P, sb, lm, scans = _scan_sequence(oracle)
    T_il = lm.T_il
    pose = lambda k: (sb.truth[k, 0:3], sb.truth[k, 3:6])  # noqa: E731
    m = lm.match_with_ref(scans[0], *pose(0))
    assert m.lines1 == [] and m.lines2 == [] and np.array_equal(m.p1, m.p2)            # no sub-map yet
    lm.add_scan(scans[0], *pose(0))
    assert lm.current_count == 1 and len(lm.ref_submap_ptr.scan_ptr.lines) == len(scans[0].lines)
    lm.add_scan(scans[0], *pose(0))                                                     # did not move: filtered
    assert lm.current_count == 1 and len(lm.key_frame) == 2
    n_ref = len(lm.ref_submap_ptr.scan_ptr.lines)
    lm.add_scan(scans[1], *pose(1))
    assert lm.current_count == 2 and len(lm.ref_submap_ptr.scan_ptr.lines) == n_ref + len(scans[1].lines)
    # a line of scan 1, as stored in the sub-map, is the same world segment as in scan 1's own frame
    from lvio2d_b200.frontend import _tf
    Tw0, Tw1 = _tf(*pose(0)) @ T_il, _tf(*pose(1)) @ T_il
    stored, own = lm.ref_submap_ptr.scan_ptr.lines[n_ref], scans[1].lines[0]
    # (to the flattening: add_line refits in the x-y plane of the sub-map's laser frame and drops z, like the reference)
    assert np.allclose((Tw0 @ np.r_[stored.p1, 1.0])[:2], (Tw1 @ np.r_[own.p1, 1.0])[:2], atol=1e-3)
    lm.add_scan(scans[2], *pose(2))
    assert lm.current_count == 3 and lm.spawnning_ref_submap_ptr is not None            # n/2 = 3: spawning sub-map created
    assert len(lm.spawnning_ref_submap_ptr.scan_ptr.lines) == len(scans[2].lines)
    match = lm.match_with_ref(scans[3], *pose(3))
    assert len(match.lines1) == len(match.lines2) >= 5
    Tw3 = _tf(*pose(3)) @ T_il
    for l1, l2 in zip(match.lines1, match.lines2):
        a, b = (Tw0 @ np.r_[l1.p1, 1.0])[:2], (Tw0 @ np.r_[l1.p2, 1.0])[:2]
        u = (b - a) / np.linalg.norm(b - a)
        for q in (l2.p1, l2.p2):
            w = (Tw3 @ np.r_[q, 1.0])[:2] - a
            assert abs(w[0] * u[1] - w[1] * u[0]) < 0.1                                  # point-to-line distance in the world
    for k in (3, 4, 5):
        lm.add_scan(scans[k], *pose(k))
    assert lm.current_count == 3 and np.array_equal(lm.ref_submap_ptr.current_p, sb.truth[2, 0:3])   # rotated at n = 6
    assert np.array_equal(lm.spawnning_ref_submap_ptr.current_p, sb.truth[5, 0:3])
    assert lm.pop_scan().scan_ptr is scans[0] and len(lm.key_frame) == 6


def _full_pipeline(oracle, backend, solver, n=8, seed=11):
    """trajectory::add_sensor_data + do_tracking for a short sequence, every step through the mirrored interfaces:
    message -> Laser -> spawn_scan -> match_with_ref -> frame_info -> solver::solve -> add_scan."""
    P = solver.params
    sb = L.synth.make_batch(1, seed, n_frames=n, beams=BEAMS, fov_deg=FOV, n_segments=12, frame_dt=0.3)
    hb = oracle.preintegrate_batch(P, sb)
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)
    lm = LaserManager(backend, L.corridor_line_params(), max_lines=160, params=P, ref_n_accumulation=100)
    a0, da = np.float32(math.radians(-FOV / 2.0)), np.float32(math.radians(FOV) / (BEAMS - 1))
    frames, out, n_pairs = [], [], []
    for k in range(n):
        pts = sb.points[sb.point_offset[k]:sb.point_offset[k + 1]]
        laser = Laser(backend, np.linalg.norm(pts, axis=1).astype(np.float32), a0, da, np.float32(0.0), 100.0 + 0.3 * k)
        scan = lm.spawn_scan(laser)
        s = sb.truth[0] if k == 0 else sb.states[k]
        f = FrameInfo(0.3 * k, s[0:3], s[3:6], s[6:9], s[9:15], imu[k - 1] if k else None, wheel[k - 1] if k else None)
        if k > 0:
            f.p += frames[-1].p - sb.states[k - 1, 0:3] if k > 1 else 0.0       # carry the last correction (cf. replay.run_tracking)
            match = lm.match_with_ref(scan, f.p, f.q)
            n_pairs.append(len(match.lines1))
            f.add_laser_match(match)
            solver.solve([frames[-1], f])
        frames.append(f)
        lm.add_scan(scan, f.p, f.q)
        out.append(np.r_[f.p, f.q])
    return sb, np.array(out), n_pairs


def test_full_front_end_tracking_pipeline_on_the_oracle(oracle):
    P = L.corridor_params(max_iters=10)
    be = oracle.OracleContext(P)
    sb, traj, n_pairs = _full_pipeline(oracle, be, Solver(P, fast_mode=True, ctx=be))
    assert min(n_pairs) >= 5
    err = np.abs(traj[1:, 0:2] - sb.truth[1:, 0:2]).max()
    guess = np.abs(sb.states[1:, 0:2] - sb.truth[1:, 0:2]).max()
    assert err < guess


@pytest.mark.gpu
def test_full_front_end_tracking_pipeline_matches_oracle(oracle):
    """The same pipeline on the CUDA library (fast_mode, 10 iterations: well-posed, see tests/test_sequence.py): every
    frame's pose within the north-star bar of the CPU path's."""
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=10)
    be = oracle.OracleContext(P)
    _, want, opairs = _full_pipeline(oracle, be, Solver(P, fast_mode=True, ctx=be))
    with Context(P) as c:
        sb, got, pairs = _full_pipeline(oracle, c, Solver(P, fast_mode=True, ctx=c))
    assert pairs == opairs
    d = np.abs(got - want)
    print(f"front-end + tracking, 8 frames: max pose difference CUDA vs oracle {d.max():.3e}")
    assert d[:, 0:3].max() <= 1e-4 and d[:, 3:6].max() <= 1e-4
