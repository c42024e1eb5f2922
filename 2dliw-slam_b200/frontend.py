"""Host-side mirror of the reference's laser front-end interface, on top of the C ABI (include/lvio2d.h).

Same names, argument meaning and call order as the reference so that tests read like the reference's own code:

  * `Laser`         = `sensor::laser` (reference src/trajectory/sensor.h:34-94): constructed from a LaserScan message
                      (`convert::laser_to_point_times`, src/utilies/common.cpp:4-40), `correct(linear, angular)` de-skews
                      in place;
  * `Scan`          = `lvio_2d::scan` (src/trajectory/laser_type.h:22-56) reduced to what the solver path consumes:
                      `lines` (p1, p2, abc), the points and index range behind every line (the line_map grid is never
                      materialised, see match_lines_kernel);
  * `LaserManager`  = `lvio_2d::laser_manager` (src/trajectory/laser_manager.h): `spawn_scan(laser)`
                      (laser_manager.cpp:350-422) and `do_match(scan1, scan2, p1, q1, p2, q2, kk)` (:244-348), the latter
                      returning the `LaserMatch` that `FrameInfo.add_laser_match` / `Solver.solve` take.

The backend is any object with `scan_to_points / extract_lines / match_lines` (`solver.Context`: the CUDA library; the
tests also pass a CPU-oracle stand-in).  The sub-map bookkeeping of `laser_manager::add_scan` (:424-496) and the
`match_with_front / back / ref` wrappers (:498-547) are host control logic in the reference and are host logic here
(plain Python on the line lists); matching against a sub-map uses the 0.05 m-sample rasterisation of
`scan::add_line(p1, p2, false)` inside `lvio2d_match_lines`.
"""
import numpy as np

from . import abi
from .params import params_T
from .solver import LaserMatch, Line
from .synth import exp_so3


def _log_so3(R):
    """rotation matrix -> rotation vector (only its norm is used, by the motion filter of add_scan)."""
    c = max(-1.0, min(1.0, (np.trace(R) - 1.0) / 2.0))
    ang = float(np.arccos(c))
    if ang < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2.0 * np.sin(ang))
    return w * ang


class Laser:
    """sensor::laser.  `ranges` float32 [n_beams]; the message fields keep their wire types (float32)."""

    def __init__(self, backend, ranges, angle_min, angle_increment, time_increment, stamp):
        self._be = backend
        self.ranges = np.ascontiguousarray(ranges, dtype=np.float32).reshape(1, -1)
        self.header = np.zeros(1, dtype=abi.SCAN_HEADER_DTYPE)
        self.header[0]["angle_min"], self.header[0]["angle_increment"] = angle_min, angle_increment
        self.header[0]["time_increment"], self.header[0]["stamp"] = time_increment, stamp
        self.time_stamp = float(stamp)
        self._convert(deskew=False)

    def _convert(self, deskew):
        cnt, pts, pz, pt = self._be.scan_to_points(self.ranges, self.header, deskew=deskew, want_times=True)
        n = int(cnt[0])
        self.points = np.concatenate([pts[0, :n], pz[0, :n, None]], axis=1)   # points_ptr  [n][3]
        self.times = pt[0, :n].copy()                                          # times_ptr

    def correct(self, linear, angular):
        """sensor::laser::correct: p <- make_tf(dt * linear, dt * angular) * p for every kept point."""
        self.header[0]["linear"], self.header[0]["angular"] = np.asarray(linear, float), np.asarray(angular, float)
        self._convert(deskew=True)


class ScanLine(Line):
    """lvio_2d::line with the fitted (a, b, c) and the index range of the scan points it was fitted to."""

    def __init__(self, p1, p2, abc, index1, index2):
        super().__init__(p1, p2)
        self.abc = np.asarray(abc, dtype=np.float64)
        self.len = float(np.linalg.norm(self.p1 - self.p2))
        self.index1, self.index2 = int(index1), int(index2)


class Scan:
    """lvio_2d::scan: the ordered `lines` of one scan, its points and its time (first kept beam)."""

    def __init__(self, time, points, lines):
        self.time, self.points, self.lines = float(time), points, lines


class LaserSubmap:
    """lvio_2d::laser_submap (laser_type.h): a scan and the IMU pose it is expressed under."""

    def __init__(self, scan, current_p, current_q):
        self.scan_ptr = scan
        self.current_p, self.current_q = np.array(current_p, dtype=np.float64), np.array(current_q, dtype=np.float64)


def _tf(p, q):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = exp_so3(np.asarray(q, dtype=np.float64)), p
    return T


def _inv_iso(T):
    """Eigen::Transform<., Isometry>::inverse(): (R^T, -R^T t) — NOT a general inverse.  The extrinsics come out of
    lie::normalize_tf without a quaternion normalisation (src/utilies/common.h:183-189), so their rotation block is
    orthonormal only to the ~1e-7 of the yaml's digits, and the reference's results carry exactly this transpose."""
    out = np.eye(4)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -(T[:3, :3].T @ T[:3, 3])
    return out


class LaserManager:
    """lvio_2d::laser_manager: spawn_scan and do_match on the device entry points; add_scan / match_with_* / pop_scan as
    the reference's host bookkeeping (key-frame deque, reference sub-map and the one being spawned)."""

    def __init__(self, backend, line_params, max_lines=256, params=None, ref_motion_filter_p=0.01, ref_motion_filter_q=0.01,
                 ref_n_accumulation=100, device_submap=False, line_cap=16384):
        self._be, self.line_params, self.max_lines = backend, line_params, int(max_lines)
        # device_submap: the reference sub-map and the spawning one live in device memory (lvio2d_submap_*), add_scan and
        # match_with_ref become one kernel each and ref_submap_ptr / spawnning_ref_submap_ptr stay None on the host
        self._dev = backend.submap(line_params, 1, line_cap, ref_motion_filter_p, ref_motion_filter_q, ref_n_accumulation) if device_submap else None
        # config/corridor.yaml:120-122
        self.ref_motion_filter_p, self.ref_motion_filter_q = ref_motion_filter_p, ref_motion_filter_q
        self.ref_n_accumulation = int(ref_n_accumulation)
        self.T_il = np.eye(4)
        if params is not None:
            self.T_il[:3, :] = params_T(params, "T_imu_to_laser")
        self.key_frame = []
        self.ref_submap_ptr = None
        self.spawnning_ref_submap_ptr = None
        self.last_add_tf = np.eye(4)
        self.current_count = 0

    # ---- scan::add_line(p1, p2, false) (laser_manager.cpp:215-223 -> :137-154, :197-213): the three fake points p1, mid,
    # p2 are collinear in x, y, so the least-squares refit returns the same 2-D line and create_line's projections return
    # the end points with z = 0 (to rounding; tests/test_ref_frontend.py measures 1e-15 against the reference text).
    # What remains are the filters: the 3-D distance of the fake points from that z = 0 line (= |z|) against
    # line_max_dis, the length against line_min_len, and the line only enters `lines` when one of its 0.05 m samples
    # falls on a valid cell of the sub-map grid.
    def _add_line(self, scan, p1, p2):
        p1, p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)
        lp = self.line_params
        if max(abs(p1[2]), abs(p2[2]), abs(0.5 * (p1[2] + p2[2]))) > lp.line_max_dis:
            return
        p1, p2 = np.array([p1[0], p1[1], 0.0]), np.array([p2[0], p2[1], 0.0])
        ln = float(np.linalg.norm(p1 - p2))
        if ln < lp.line_min_len:
            return
        w, h = int(lp.w_laser_each_scan / lp.laser_resolution + 1), int(lp.h_laser_each_scan / lp.laser_resolution + 1)
        d = p2 - p1
        unit = d / ln
        tr, hit = 0.0, False
        while tr <= ln and not hit:
            q = p1 + unit * tr
            c, r = int(q[0] / lp.laser_resolution + w // 2), int(q[1] / lp.laser_resolution + h // 2)
            hit = 0 <= r < h and 0 <= c < w
            tr += 0.05
        if not hit:
            return
        nrm = np.array([-d[1], d[0]]) / np.linalg.norm(d[:2])
        scan.lines.append(ScanLine(p1, p2, [nrm[0], nrm[1], -float(nrm @ p1[:2])], 0, -1))

    def _empty_submap(self, current_p, current_q):
        return LaserSubmap(Scan(0.0, None, []), current_p, current_q)

    def add_scan(self, scan, current_p, current_q):
        """laser_manager::add_scan (laser_manager.cpp:424-496)."""
        self.key_frame.append(LaserSubmap(scan, current_p, current_q))
        if self._dev is not None:
            n, lines, _ = self._pack(scan, max(1, len(scan.lines)))
            self._dev.add_scan(n, lines, np.concatenate([current_p, current_q]))
            return
        current_tf = _tf(current_p, current_q)
        if self.ref_submap_ptr is not None:
            d = _inv_iso(self.last_add_tf) @ current_tf
            dq = _log_so3(d[:3, :3])
            if np.linalg.norm(d[:3, 3]) < self.ref_motion_filter_p and np.linalg.norm(dq) < self.ref_motion_filter_q:
                return
        else:
            self.ref_submap_ptr = self._empty_submap(current_p, current_q)
            self.last_add_tf = current_tf
            self.current_count = 1
            for l in scan.lines:
                self._add_line(self.ref_submap_ptr.scan_ptr, l.p1, l.p2)
            return
        for sub in (self.ref_submap_ptr, self.spawnning_ref_submap_ptr):
            if sub is None:
                continue
            T = _inv_iso(self.T_il) @ (_inv_iso(_tf(sub.current_p, sub.current_q)) @ current_tf) @ self.T_il
            for l in scan.lines:
                self._add_line(sub.scan_ptr, T[:3, :3] @ l.p1 + T[:3, 3], T[:3, :3] @ l.p2 + T[:3, 3])
        self.current_count += 1
        if self.spawnning_ref_submap_ptr is None and self.current_count == self.ref_n_accumulation // 2:
            self.spawnning_ref_submap_ptr = self._empty_submap(current_p, current_q)
            self.last_add_tf = current_tf
            for l in scan.lines:
                self._add_line(self.spawnning_ref_submap_ptr.scan_ptr, l.p1, l.p2)
        if self.current_count == self.ref_n_accumulation:
            self.ref_submap_ptr = self.spawnning_ref_submap_ptr
            self.spawnning_ref_submap_ptr = self._empty_submap(current_p, current_q)
            self.last_add_tf = current_tf
            for l in scan.lines:
                self._add_line(self.spawnning_ref_submap_ptr.scan_ptr, l.p1, l.p2)
            self.current_count = self.ref_n_accumulation // 2
        self.last_add_tf = current_tf

    def _no_match(self, scan, current_p, current_q):
        m = LaserMatch([], [], current_p, current_q)
        m.p2, m.q2 = np.array(current_p, dtype=np.float64), np.array(current_q, dtype=np.float64)
        m.scan2 = scan
        return m

    def match_with_front(self, scan, current_p, current_q):
        if not self.key_frame:
            return self._no_match(scan, current_p, current_q)
        kf = self.key_frame[0]
        return self.do_match(kf.scan_ptr, scan, kf.current_p, kf.current_q, current_p, current_q)

    def match_with_back(self, scan, current_p, current_q):
        if not self.key_frame:
            return self._no_match(scan, current_p, current_q)
        kf = self.key_frame[-1]
        return self.do_match(kf.scan_ptr, scan, kf.current_p, kf.current_q, current_p, current_q)

    def match_with_ref(self, scan, current_p, current_q):
        if self._dev is not None:
            meta = self._dev.get(0, want_lines=False)[0]
            if not meta[0, 0]:
                return self._no_match(scan, current_p, current_q)
            n2, l2, _ = self._pack(scan, max(1, len(scan.lines)))
            nm, m, l1, pose = self._dev.match(n2, l2, np.concatenate([current_p, current_q]), want_lines1=True)
            pairs = m[0, :int(nm[0])]
            match = LaserMatch([Line([*l1[0, k, 0:2], 0.0], [*l1[0, k, 2:4], 0.0]) for k in range(len(pairs))], [scan.lines[int(i)] for _, i in pairs],
                               pose[0, 0:3], pose[0, 3:6])
            match.p2, match.q2 = np.array(current_p, dtype=np.float64), np.array(current_q, dtype=np.float64)
            match.scan2 = scan
            match.index_pairs = pairs.copy()
            return match
        if self.ref_submap_ptr is None:
            return self._no_match(scan, current_p, current_q)
        r = self.ref_submap_ptr
        return self.do_match(r.scan_ptr, scan, r.current_p, r.current_q, current_p, current_q)

    def pop_scan(self):
        return self.key_frame.pop(0) if self.key_frame else None

    def spawn_scan(self, laser):
        pts = laser.points
        off = np.array([0, len(pts)], dtype=np.int64)
        n, lines, abc, rng = self._be.extract_lines(self.line_params, off, pts[:, 0:2], max_lines=self.max_lines, point_z=pts[:, 2])
        k = min(int(n[0]), self.max_lines)
        out = [ScanLine([*lines[0, j, 0:2], 0.0], [*lines[0, j, 2:4], 0.0], abc[0, j], rng[0, j, 0], rng[0, j, 1]) for j in range(k)]
        return Scan(laser.times[0] if len(laser.times) else laser.time_stamp, pts, out)

    @staticmethod
    def _pack(scan, max_lines):
        n = np.array([len(scan.lines)], dtype=np.int32)
        lines = np.zeros((1, max_lines, 4))
        rng = np.zeros((1, max_lines, 2), dtype=np.int32)
        for j, l in enumerate(scan.lines):
            lines[0, j] = [l.p1[0], l.p1[1], l.p2[0], l.p2[1]]
            rng[0, j] = [l.index1, l.index2]
        return n, lines, rng

    def do_match(self, scan1, scan2, p1, q1, p2, q2, kk=0):
        """laser_manager::do_match: the LaserMatch of scan2 (current) against scan1 (reference)."""
        m1, m2 = max(1, len(scan1.lines)), max(1, len(scan2.lines))
        n1, l1, r1 = self._pack(scan1, m1)
        n2, l2, _ = self._pack(scan2, m2)
        pose1, pose2 = np.concatenate([p1, q1]).reshape(1, 6), np.concatenate([p2, q2]).reshape(1, 6)
        if scan1.points is not None:     # a scan from spawn_scan: its lines cover the cells of their own points
            kw = dict(point_offset1=np.array([0, len(scan1.points)], dtype=np.int64),
                      points1=np.ascontiguousarray(scan1.points[:, 0:2]), index_range1=r1)
        else:                            # a sub-map built by add_line(p1, p2, false): cells sampled along the segments
            kw = {}
        nm, m = self._be.match_lines(self.line_params, n1, l1, n2, l2, pose1, pose2, kk=kk, **kw)
        lines1 = [scan1.lines[int(j)] for j, _ in m[0, :int(nm[0])]]
        lines2 = [scan2.lines[int(i)] for _, i in m[0, :int(nm[0])]]
        match = LaserMatch(lines1, lines2, p1, q1)
        match.p2, match.q2 = np.array(p2, dtype=np.float64), np.array(q2, dtype=np.float64)
        match.scan2 = scan2
        return match
