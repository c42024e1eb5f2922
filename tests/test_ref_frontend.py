"""Pins the laser front-end oracles (oracle/scan_points.hpp, laser_lines.hpp, laser_match.hpp) and the host mirror of the
sub-map bookkeeping (lvio2d_b200.frontend.LaserManager) to the REFERENCE'S OWN src/utilies/common.cpp,
src/trajectory/sensor.h and src/trajectory/laser_manager.cpp, compiled unmodified into oracle/_ref/libref.so against the
stub Eigen tree (tests/ref_lib.py).  Integer results (point counts, line counts, match pairs) must be equal; point
coordinates to 1e-12, fitted line end points to 1e-10 (the stub's JacobiSVD and the oracle's both differ from Eigen's in
rounding only; measured 1.2e-12)."""
import math

import numpy as np
import pytest

import lvio2d_b200 as L
import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not built and /root/reference absent")


def test_scan_to_points_matches_reference_text(oracle):
    """convert::laser_to_point_times (float32 angle arithmetic, NaN / inf / < 0.1 m dropped, 1 cm thinning against the
    last KEPT point) + sensor::laser::correct."""
    rg, hd = L.synth.make_range_batch(6, 3, beams=1081)
    for deskew in (False, True):
        cnt, pts, pz, pt = oracle.scan_to_points(rg, hd, deskew=deskew)
        for k in range(len(rg)):
            rp, rt = ref_lib.scan_to_points(rg[k], hd[k], deskew=deskew)
            assert len(rp) == cnt[k]
            assert np.array_equal(rt, pt[k, :cnt[k]])
            assert np.abs(rp[:, 0:2] - pts[k, :cnt[k]]).max() < 1e-12
            assert np.abs(rp[:, 2] - pz[k, :cnt[k]]).max() < 1e-12
    # edge cases: nothing valid, everything within 1 cm of the first point
    bad = np.full(64, np.nan, np.float32)
    bad[::2] = 0.05
    assert len(ref_lib.scan_to_points(bad, hd[0], False)[0]) == 0 and oracle.scan_to_points(bad[None], hd[:1], False)[0][0] == 0
    near = np.full(64, 0.2, np.float32)
    h = hd[0].copy()
    h["angle_increment"] = np.float32(1e-4)
    assert len(ref_lib.scan_to_points(near, h, False)[0]) == oracle.scan_to_points(near[None], np.array([h]), False)[0][0]


def lines_by_scan(off, pts, lp, oracle, max_lines=160):
    n, lines, abc, rng = oracle.extract_lines(lp, off, pts, max_lines=max_lines)
    return n, lines, abc, rng


def test_spawn_scan_matches_reference_text(oracle):
    """laser_manager::spawn_scan + scan::add_line(points, i1, i2): continuity split, corner response, non-maximum
    suppression, tolerance-angle merge, least-squares fit through JacobiSVD, create_line, the three filters."""
    lp = L.corridor_line_params()
    off, pts = L.synth.make_scan_batch(10, 11, beams=1081, range_sigma=0.004)
    n, lines, abc, rng = lines_by_scan(off, pts, lp, oracle)
    total = 0
    for k in range(len(off) - 1):
        p = pts[off[k]:off[k + 1]]
        sc = ref_lib.RefScan.from_points(np.c_[p, np.zeros(len(p))])
        rl, _ = sc.lines()
        assert len(rl) == n[k], (k, len(rl), n[k])
        got = lines[k, :n[k]]
        assert np.abs(rl[:, [0, 1, 3, 4]] - got).max() < 1e-10, k     # measured 1.2e-12
        # abc = the smallest right singular vector: defined up to sign
        for j in range(n[k]):
            assert min(np.abs(rl[j, 6:9] - abc[k, j]).max(), np.abs(rl[j, 6:9] + abc[k, j]).max()) < 1e-10
        total += n[k]
    assert total > 150


def test_do_match_matches_reference_text(oracle):
    """laser_manager::do_match with scan1 built from points (line_map filled from the scan points) and from lines
    (the sub-map flavour, 0.05 m samples): pairs as index pairs, in order."""
    P, lp = L.corridor_params(), L.corridor_line_params()
    off, pts = L.synth.make_scan_batch(6, 23, beams=1081, range_sigma=0.004)
    n, lines, abc, rng = lines_by_scan(off, pts, lp, oracle)
    rgen = np.random.default_rng(5)
    pairs_seen = 0
    for k in range(len(off) - 1):
        p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
        s1 = ref_lib.RefScan.from_points(p3)
        # scan 2 = the same scan seen from a slightly moved pose: transform the points into the moved laser frame
        pose1 = np.array([1.0, 2.0, 0.0, *ref_lib.log_SO3(np.eye(3))])
        from scipy.spatial.transform import Rotation
        T_il = np.array(list(P.T_imu_to_laser)).reshape(3, 4)
        base = Rotation.from_matrix(T_il[:, :3].T)          # imu orientation that levels the laser
        pose1 = np.r_[rgen.uniform(-2, 2, 2), 0.0, base.as_rotvec()]
        dyaw, dxy = rgen.normal(0, 0.02), rgen.normal(0, 0.03, 2)
        R1 = base.as_matrix()
        # move the LASER by (dxy, dyaw) in its own plane
        Rl = Rotation.from_euler("z", dyaw).as_matrix()
        Rwl1 = R1 @ T_il[:, :3]
        twl1 = R1 @ T_il[:, 3] + pose1[0:3]
        Rwl2 = Rwl1 @ Rl
        twl2 = twl1 + Rwl1 @ np.r_[dxy, 0.0]
        R2 = Rwl2 @ T_il[:, :3].T
        pose2 = np.r_[twl2 - R2 @ T_il[:, 3], Rotation.from_matrix(R2).as_rotvec()]
        q3 = (Rl.T @ (p3 - np.r_[dxy, 0.0]).T).T
        s2 = ref_lib.RefScan.from_points(q3)
        l2, _ = s2.lines()
        n2, lines2, _, _ = oracle.extract_lines(lp, np.array([0, len(q3)], np.int64), q3[:, :2], max_lines=160)
        assert n2[0] == len(l2)
        for kk in (0, 1):
            want = ref_lib.do_match(s1, s2, pose1, pose2, kk)
            nm, match = oracle.match_lines(P, lp, n[k:k + 1], lines[k:k + 1], n2, lines2, pose1[None], pose2[None], kk=kk,
                                           point_offset1=np.array([0, len(p3)], np.int64), points1=p3[:, :2], index_range1=rng[k:k + 1])
            assert nm[0] == len(want) and np.array_equal(match[0, :nm[0]], want), (k, kk)
            pairs_seen += len(want)
        # sub-map flavour: scan 1 rebuilt from its lines by add_line(p1, p2, false)
        l1, _ = s1.lines()
        sm = ref_lib.RefScan.from_lines(l1[:, 0:6])
        sl, _ = sm.lines()
        want = ref_lib.do_match(sm, s2, pose1, pose2, 0)
        n1s = np.array([len(sl)], np.int32)
        l1s = np.zeros((1, 160, 4))
        l1s[0, :len(sl)] = sl[:, [0, 1, 3, 4]]
        nm, match = oracle.match_lines(P, lp, n1s, l1s, n2, lines2, pose1[None], pose2[None], kk=0)
        assert nm[0] == len(want) and np.array_equal(match[0, :nm[0]], want), k
    assert pairs_seen > 60


def test_add_scan_submap_bookkeeping_matches_reference_text(oracle):
    """laser_manager::add_scan (laser_manager.cpp:424-496): motion filter, the reference sub-map and the spawning one,
    the hand-over at ref_n_accumulation — lvio2d_b200.frontend.LaserManager against the reference class, frame by frame.
    corridor.yaml's ref_n_accumulation is 100: the sequence below is long enough to roll the buffers twice."""
    from scipy.spatial.transform import Rotation

    from types import SimpleNamespace

    from lvio2d_b200.frontend import LaserManager

    P, lp = L.corridor_params(), L.corridor_line_params()
    T_il = np.array(list(P.T_imu_to_laser)).reshape(3, 4)
    base = Rotation.from_matrix(T_il[:, :3].T)
    off, pts = L.synth.make_scan_batch(4, 31, beams=721, range_sigma=0.004)
    be = oracle.OracleContext(P)
    mine = LaserManager(be, lp, max_lines=160, params=P, ref_n_accumulation=ref_lib.REF_N_ACCUMULATION)
    ref = ref_lib.RefLaserManager()
    rgen = np.random.default_rng(8)
    xy, yaw = np.zeros(2), 0.0
    rolled = matched = 0
    last_ref_pose = None
    for f in range(260):
        k = f % 4
        p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
        rs = ref_lib.RefScan.from_points(p3)
        ms = mine.spawn_scan(SimpleNamespace(points=p3, times=np.zeros(1), time_stamp=0.0))
        # every 5th frame does not move: the motion filter must drop it on both sides
        if f % 5:
            xy = xy + rgen.normal(0, 0.02, 2) + np.array([0.015, 0.0])
            yaw += rgen.normal(0, 0.01)
        Rwi = Rotation.from_euler("z", yaw).as_matrix() @ base.as_matrix()
        pose = np.r_[xy, 0.0, Rotation.from_matrix(Rwi).as_rotvec()]
        # trajectory.cpp's call order: match_with_ref against the sub-map as it stands, THEN add_scan (a scan matched against
        # a sub-map that already holds its own copy is degenerate: acos of a dot product that rounds above 1 is NaN and
        # the partner choice becomes a property of the last bit)
        if f % 7 == 3:
            want, rpose = ref.match_with_ref(rs, pose)
            m = mine.match_with_ref(ms, pose[0:3], pose[3:6])
            got = np.array([[mine.ref_submap_ptr.scan_ptr.lines.index(a), ms.lines.index(b)] for a, b in zip(m.lines1, m.lines2)], np.int32).reshape(-1, 2)
            assert np.array_equal(got, want), f
            assert np.abs(np.r_[m.p1, m.q1] - rpose).max() < 1e-15
            matched += len(want)
        ref.add_scan(rs, pose)
        mine.add_scan(ms, pose[0:3], pose[3:6])
        for which, sub in ((0, mine.ref_submap_ptr), (1, mine.spawnning_ref_submap_ptr)):
            r = ref.submap(which)
            if r is None:
                assert sub is None, (f, which)
                continue
            rpose, rlines, rcount = r
            assert sub is not None and rcount == mine.current_count, (f, which, rcount, mine.current_count)
            assert np.abs(np.r_[sub.current_p, sub.current_q] - rpose).max() < 1e-15
            ml = np.array([[*l.p1, *l.p2] for l in sub.scan_ptr.lines]).reshape(-1, 6)
            assert len(ml) == len(rlines), (f, which, len(ml), len(rlines))
            if len(ml):
                assert np.abs(ml - rlines).max() < 1e-9, (f, which)
        r0 = ref.submap(0)
        if last_ref_pose is not None and np.abs(r0[0] - last_ref_pose).max() > 0:
            rolled += 1
        last_ref_pose = r0[0].copy()
    assert rolled >= 2 and matched > 100
