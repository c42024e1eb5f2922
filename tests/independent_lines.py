"""Independent re-derivation of the scan -> line segments step (laser_manager::spawn_scan, reference
src/trajectory/laser_manager.cpp:350-422 with scan::add_line :137-154) used to cross-check the C++ oracle:
vectorised numpy for the continuity split / corner response / non-maximum suppression, numpy.linalg.svd for the fit,
the cross-product form of the point-line distance.  TEST INFRASTRUCTURE."""
import math

import numpy as np

EPS = 0.0008


def _cos_at(p, j, i, k):
    a, b = p[i] - p[j], p[k] - p[j]
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    if na < EPS or nb < EPS:
        return -1.0
    return float(np.dot(a / na, b / nb))


def _fit(p):
    A = np.c_[p, np.ones(len(p))]
    v = np.linalg.svd(A, full_matrices=False)[2][2]
    k = int(np.argmax(np.abs(v)))
    return v if v[k] > 0 else -v


def _try_line(lp, p, i1, i2):
    if i2 - i1 < 2:
        return None
    a, b, c = abc = _fit(p[i1:i2 + 1])
    if abs(b) < 0.5:
        q1, q2 = np.array([-c / a, 0.0]), np.array([(-c - b) / a, 1.0])
    else:
        q1, q2 = np.array([0.0, -c / b]), np.array([1.0, (-c - a) / b])
    u = (q2 - q1) / np.linalg.norm(q2 - q1)
    rel = p[i1:i2 + 1] - q2
    err = np.abs(rel[:, 0] * u[1] - rel[:, 1] * u[0]).max()
    e1 = q1 + np.dot(p[i1] - q1, u) * u
    e2 = q1 + np.dot(p[i2] - q1, u) * u
    if err > lp.line_max_dis or np.linalg.norm(e1 - e2) < lp.line_min_len:
        return None
    w = int(lp.w_laser_each_scan / lp.laser_resolution + 1)
    h = int(lp.h_laser_each_scan / lp.laser_resolution + 1)
    cc = np.trunc(p[i1:i2 + 1, 0] / lp.laser_resolution + w // 2)
    rr = np.trunc(p[i1:i2 + 1, 1] / lp.laser_resolution + h // 2)
    if not np.any((rr >= 0) & (rr < h) & (cc >= 0) & (cc < w)):
        return None
    return (i1, i2, e1, e2, abc)


def spawn_scan(lp, pts):
    p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    n = len(p)
    out = []
    if n == 0:
        return out
    gaps = np.linalg.norm(np.diff(p, axis=0), axis=1)
    brk = np.nonzero(~(gaps <= lp.line_continuous_threshold))[0] + 1
    starts = np.r_[0, brk]
    ends = np.r_[brk - 1, n - 1]
    tol = lp.line_max_tolerance_angle_deg / 180.0 * math.pi
    for s, e in zip(starts, ends):
        s, e = int(s), int(e)
        resp = np.full(n, -1.0)
        for i in range(s + 1, e):
            resp[i] = _cos_at(p, i, max(i - 3, s), min(i + 3, e))
        cand = [s]
        i = s + 1
        while i <= e - 1:
            lo, hi = max(i - 3, s + 1), min(i + 3, e - 1)
            window = np.r_[resp[lo:i], resp[i + 1:hi + 1]]
            if not np.any(window >= resp[i]):
                cand.append(i)
                i += 3
            i += 1
        cand.append(e)
        last = 0
        for k in range(1, len(cand) - 1):
            c = _cos_at(p, cand[k], cand[last], cand[k + 1])
            ang = math.acos(c) if -1.0 <= c <= 1.0 else float("nan")
            if abs(ang) < tol:
                ln = _try_line(lp, p, cand[last], cand[k])
                if ln:
                    out.append(ln)
                last = k
        ln = _try_line(lp, p, cand[last], cand[-1])
        if ln:
            out.append(ln)
    return out
