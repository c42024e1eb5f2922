"""The association oracle (oracle/laser_match.hpp = laser_manager::do_match, reference
src/trajectory/laser_manager.cpp:244-348) against an independent Python restatement (scipy rotations, dict grid,
cross-product distances) and hand-built cases."""
import math

import numpy as np
from scipy.spatial.transform import Rotation

import lvio2d_b200 as L
from lvio2d_b200.params import params_T


def _iso(pose, T):
    R = Rotation.from_rotvec(pose[3:6]).as_matrix()
    return R @ T[:, :3], R @ T[:, 3] + pose[0:3]


def restate(P, lp, n1, l1, n2, l2, pose1, pose2, kk, pts=None, rng=None):
    T_il = params_T(P, "T_imu_to_laser")
    w = int(lp.w_laser_each_scan / lp.laser_resolution + 1)
    h = int(lp.h_laser_each_scan / lp.laser_resolution + 1)
    cell = lambda x, y: (int(y / lp.laser_resolution + h // 2), int(x / lp.laser_resolution + w // 2))  # noqa: E731
    ok = lambda r, c: 0 <= r < h and 0 <= c < w  # noqa: E731
    grid = {}
    for j in range(n1):
        if pts is not None:
            xs = [pts[i] for i in range(rng[j, 0], rng[j, 1] + 1)]
        else:
            a, b = l1[j, :2], l1[j, 2:]
            ln = np.linalg.norm(a - b)
            u = (b - a) / ln
            xs, tr = [], 0.0
            while tr <= ln:
                xs.append(a + u * tr)
                tr += 0.05
        for x in xs:
            rc = cell(x[0], x[1])
            if ok(*rc):
                lst = grid.setdefault(rc, [])
                if not lst or lst[-1] != j:
                    lst.append(j)
    R1, t1 = _iso(pose1, T_il)
    R2, t2 = _iso(pose2, T_il)
    R12, t12 = R1.T @ R2, R1.T @ (t2 - t1)
    tf = lambda p: R12 @ np.array([p[0], p[1], 0.0]) + t12  # noqa: E731
    ret = []
    for i in range(n2):
        a, b = l2[i, :2], l2[i, 2:]
        m = tf((a + b) / 2)
        r, c = cell(m[0], m[1])
        cand = []
        for dr in range(-1 - kk, 2 + kk):
            for dc in range(-1 - kk, 2 + kk):
                if ok(r + dr, c + dc):
                    cand += grid.get((r + dr, c + dc), [])
        if not cand:
            continue
        v2 = tf(b) - tf(a)
        v2 = v2 / np.linalg.norm(v2)
        best, best_a = None, 2 * math.pi
        for j in cand:
            v1 = np.r_[l1[j, 2:] - l1[j, :2], 0.0]
            ang = math.acos(min(1.0, abs(float(np.dot(v1 / np.linalg.norm(v1), v2)))))
            if ang < best_a:
                best, best_a = j, ang
        if math.degrees(best_a) > 10:
            continue
        ret.append((best, i))

    def dist(p, l):
        u = np.r_[l[2:] - l[:2], 0.0]
        u = u / np.linalg.norm(u)
        return np.linalg.norm(np.cross(p - np.r_[l[2:], 0.0], u))

    ds = [0.5 * (dist(tf(l2[i, :2]), l1[j]) + dist(tf(l2[i, 2:]), l1[j])) for j, i in ret]
    aver = sum(ds) / len(ds) if ds else 0.0
    return [pr for pr, dd in zip(ret, ds) if dd < aver * 1.2]


def test_match_oracle_matches_independent_restatement(oracle):
    P, lp = L.corridor_params(), L.corridor_line_params()
    sb = L.synth.make_batch(6, 9, n_frames=2, beams=1081, n_segments=12, frame_dt=0.3)
    off = sb.point_offset
    n, lines, _, rng = oracle.extract_lines(lp, off, sb.points, max_lines=128)
    i1, i2 = np.arange(0, 12, 2), np.arange(1, 12, 2)
    pose1, pose2 = sb.truth[i1, :6], sb.truth[i2, :6] + np.random.default_rng(1).normal(0, 0.01, (6, 6))
    cnt1 = np.diff(off)[i1].astype(np.int32)
    for kk in (0, 1):
        nm, m = oracle.match_lines(P, lp, n[i1], lines[i1], n[i2], lines[i2], pose1, pose2, kk, off[i1], sb.points, rng[i1], cnt1)
        nm2, m2 = oracle.match_lines(P, lp, n[i1], lines[i1], n[i2], lines[i2], pose1, pose2, kk)
        for p in range(6):
            pts = sb.points[off[i1[p]]:off[i1[p] + 1]]
            want = restate(P, lp, n[i1[p]], lines[i1[p]], n[i2[p]], lines[i2[p]], pose1[p], pose2[p], kk, pts, rng[i1[p]])
            assert [tuple(x) for x in m[p, :nm[p]]] == want
            want = restate(P, lp, n[i1[p]], lines[i1[p]], n[i2[p]], lines[i2[p]], pose1[p], pose2[p], kk)
            assert [tuple(x) for x in m2[p, :nm2[p]]] == want
        assert nm.sum() > 60


def test_match_known_answers(oracle):
    """Identity relative pose: every line of scan 2 that also exists in scan 1 is matched with itself; a line rotated by
    more than 10 degrees is rejected; a far-away line finds no candidates."""
    P, lp = L.corridor_params(), L.corridor_line_params()
    l1 = np.zeros((1, 4, 4))
    l1[0, 0] = [1.0, -1.0, 1.3, 2.0]
    l1[0, 1] = [-2.0, 0.5, 0.0, 1.7]
    l1[0, 2] = [3.0, 3.0, 5.0, 3.4]
    l2 = np.zeros((1, 4, 4))
    l2[0, 0] = l1[0, 1]
    l2[0, 1] = l1[0, 0]
    c, s = math.cos(0.3), math.sin(0.3)
    mid = (l1[0, 2, :2] + l1[0, 2, 2:]) / 2
    half = (l1[0, 2, 2:] - l1[0, 2, :2]) / 2
    rot = np.array([c * half[0] - s * half[1], s * half[0] + c * half[1]])
    l2[0, 2] = np.r_[mid - rot, mid + rot]           # same place, turned by 17 degrees
    l2[0, 3] = [20.0, 20.0, 21.0, 20.5]              # nothing of scan 1 nearby
    pose = np.array([[0.3, -0.2, 0.0, 0.0, 0.0, 0.4]])
    nm, m = oracle.match_lines(P, lp, [3], l1, [4], l2, pose, pose)
    assert nm[0] == 2 and m[0, :2].tolist() == [[1, 0], [0, 1]]
