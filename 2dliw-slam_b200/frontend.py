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
tests also pass a CPU-oracle stand-in).  The sub-map bookkeeping of `laser_manager::add_scan` (:424-496) is host logic
that stays with the caller.
"""
import numpy as np

from . import abi
from .solver import LaserMatch, Line


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


class LaserManager:
    """lvio_2d::laser_manager restricted to spawn_scan and do_match."""

    def __init__(self, backend, line_params, max_lines=256):
        self._be, self.line_params, self.max_lines = backend, line_params, int(max_lines)

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
        nm, m = self._be.match_lines(self.line_params, n1, l1, n2, l2, pose1, pose2, kk=kk,
                                     point_offset1=np.array([0, len(scan1.points)], dtype=np.int64),
                                     points1=np.ascontiguousarray(scan1.points[:, 0:2]), index_range1=r1)
        lines1 = [scan1.lines[int(j)] for j, _ in m[0, :int(nm[0])]]
        lines2 = [scan2.lines[int(i)] for _, i in m[0, :int(nm[0])]]
        match = LaserMatch(lines1, lines2, p1, q1)
        match.p2, match.q2 = np.array(p2, dtype=np.float64), np.array(q2, dtype=np.float64)
        match.scan2 = scan2
        return match
