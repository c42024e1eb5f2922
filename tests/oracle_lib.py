"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE (the CPU restatement of the reference path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The library is built by `make -C oracle` (also done by __graft_entry__.build()).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from lvio2d_b200 import abi  # noqa: E402

_LIB_PATH = os.path.join(_ROOT, "oracle", "liboracle.so")
_lib = None

dp = abi.c_double_p


def _d(a):
    return a.ctypes.data_as(dp)


def build(force=False):
    srcs = ["oracle_c.cpp", "laser_lines.hpp", "scan_points.hpp", "laser_match.hpp", "pose_graph.hpp", "solver.hpp", "factors.hpp", "preint.hpp", "lie.hpp", "jet.hpp"]
    newest = max(os.path.getmtime(os.path.join(_ROOT, "oracle", s)) for s in srcs)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_dis_from_line.restype = C.c_double
        _lib.oracle_dis_from_line.argtypes = [dp, dp, dp]
        _lib.oracle_eval_laser_point.argtypes = [C.POINTER(abi.Params), dp, dp, dp, C.c_double, dp, dp, dp, dp]
    return _lib


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


# ---- primitives
def exp_so3(v):
    R = np.zeros(9)
    lib().oracle_exp_so3(_d(_arr(v)), _d(R))
    return R.reshape(3, 3)


def log_SO3(R):
    v = np.zeros(3)
    lib().oracle_log_SO3(_d(_arr(R).ravel()), _d(v))
    return v


def normalize_so3(v):
    v = _arr(v).copy()
    lib().oracle_normalize_so3(_d(v))
    return v


def so3_plus(theta, delta):
    out = np.zeros(3)
    lib().oracle_so3_plus(_d(_arr(theta)), _d(_arr(delta)), _d(out))
    return out


def dis_from_line(p, p1, p2):
    return lib().oracle_dis_from_line(_d(_arr(p)), _d(_arr(p1)), _d(_arr(p2)))


# ---- factors
def eval_laser_factor(params, l1_p1, l1_p2, l2_p1, l2_p2, pose_i, pose_j):
    res, jac = np.zeros(2), np.zeros((2, 12))
    lib().oracle_eval_laser_factor(C.byref(params), _d(_arr(l1_p1)), _d(_arr(l1_p2)), _d(_arr(l2_p1)), _d(_arr(l2_p2)),
                                   _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def eval_laser_point(params, a1, a2, c, weight, pose_i, pose_j):
    res, jac = np.zeros(1), np.zeros((1, 12))
    lib().oracle_eval_laser_point(C.byref(params), _d(_arr(a1)), _d(_arr(a2)), _d(_arr(c)), float(weight),
                                  _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def eval_laser_point_analytic(params, a1, a2, c, weight, pose_i, pose_j):
    """The closed-form flavour of the same factor (cpu_baseline "analytic")."""
    res, jac = np.zeros(1), np.zeros((1, 12))
    lib().oracle_eval_laser_point_analytic.argtypes = [C.POINTER(abi.Params), dp, dp, dp, C.c_double, dp, dp, dp, dp]
    lib().oracle_eval_laser_point_analytic(C.byref(params), _d(_arr(a1)), _d(_arr(a2)), _d(_arr(c)), float(weight),
                                           _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def eval_imu_factor(params, blob, state_i, state_j):
    res, jac = np.zeros(15), np.zeros((15, 30))
    lib().oracle_eval_imu_factor(C.byref(params), _d(_arr(blob)), _d(_arr(state_i)), _d(_arr(state_j)), _d(res), _d(jac))
    return res, jac


def eval_wheel_factor(params, blob, pose_i, pose_j):
    res, jac = np.zeros(3), np.zeros((3, 12))
    lib().oracle_eval_wheel_factor(C.byref(params), _d(_arr(blob)), _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def eval_ground_factors(params, pose):
    res, jac = np.zeros(2), np.zeros((2, 6))
    lib().oracle_eval_ground_factors(C.byref(params), _d(_arr(pose)), _d(res), _d(jac))
    return res, jac


def eval_prior_factor(X0, J, state):
    res, jac = np.zeros(15), np.zeros((15, 15))
    lib().oracle_eval_prior_factor(_d(_arr(X0)), _d(_arr(J).ravel()), _d(_arr(state)), _d(res), _d(jac))
    return res, jac


# ---- preintegration
def imu_preintegrate(params, sample_offset, samples, bias0):
    off = np.ascontiguousarray(sample_offset, dtype=np.int64)
    n = off.size - 1
    out = np.zeros((n, abi.IMU_BLOB))
    rc = lib().oracle_imu_preintegrate(C.byref(params), n, off.ctypes.data_as(abi.c_int64_p), _d(_arr(samples)), _d(_arr(bias0)), _d(out))
    assert rc == 0
    return out


def wheel_preintegrate(params, step_offset, steps):
    off = np.ascontiguousarray(step_offset, dtype=np.int64)
    n = off.size - 1
    out = np.zeros((n, abi.WHEEL_BLOB))
    rc = lib().oracle_wheel_preintegrate(C.byref(params), n, off.ctypes.data_as(abi.c_int64_p), _d(_arr(steps)), _d(out))
    assert rc == 0
    return out


def preintegrate_batch(params, sb):
    """SensorBatch -> HostBatch through the ORACLE preintegrators (CPU tests only)."""
    if sb.n_frames > 1:
        imu = imu_preintegrate(params, sb.imu_offset, sb.imu_samples, sb.bias0)
        wheel = wheel_preintegrate(params, sb.wheel_offset, sb.wheel_steps)
    else:
        imu = wheel = None
    return sb.host_batch(imu, wheel)


# ---- window level
def linearize(params, hb, states=None, mode=0):
    B, dim = hb.n_windows, 15 * hb.n_frames
    H, g, cost = np.zeros((B, dim, dim)), np.zeros((B, dim)), np.zeros(B)
    s = hb.struct()
    st = None if states is None else _arr(states)
    rc = lib().oracle_linearize(C.byref(params), C.byref(s), _d(st) if st is not None else dp(), int(mode), _d(H), _d(g), _d(cost))
    assert rc == 0
    return H, g, cost


def cost(params, hb, states=None):
    out = np.zeros(hb.n_windows)
    s = hb.struct()
    st = None if states is None else _arr(states)
    lib().oracle_cost(C.byref(params), C.byref(s), _d(st) if st is not None else dp(), _d(out))
    return out


def solve(params, hb, n_threads=1, analytic=False):
    """ceres::Solve restated.  analytic=True: closed-form Jacobian of the scan-point factor instead of Jets (the faster CPU
    baseline flavour; same minimiser)."""
    B, n = hb.n_windows, hb.n_frames
    states = np.zeros((B * n, 15))
    summ = np.zeros(B, dtype=abi.SUMMARY_DTYPE)
    s = hb.struct()
    fn = lib().oracle_solve_analytic if analytic else lib().oracle_solve
    rc = fn(C.byref(params), C.byref(s), _d(states), summ.ctypes.data_as(C.c_void_p), int(n_threads))
    assert rc == 0
    return states, summ


def linear_solve_seconds(reset=False):
    """Seconds this thread's single-threaded solves spent in the linear solve (dense reduced system) since the last reset."""
    fn = lib().oracle_linear_solve_seconds
    fn.restype = C.c_double
    fn.argtypes = [C.c_int32]
    return float(fn(1 if reset else 0))


def marginalize(params, hb, states=None):
    B = hb.n_windows
    X0, J, r, dH, dg = np.zeros((B, 15)), np.zeros((B, 15, 15)), np.zeros((B, 15)), np.zeros((B, 15, 15)), np.zeros((B, 15))
    s = hb.struct()
    st = None if states is None else _arr(states)
    rc = lib().oracle_marginalize(C.byref(params), C.byref(s), _d(st) if st is not None else dp(), _d(X0), _d(J), _d(r), _d(dH), _d(dg))
    assert rc == 0
    return X0, J, r, dH, dg


def extract_lines(lp, point_offset, points, max_lines=256, point_count=None, point_z=None):
    """laser_manager::spawn_scan for a batch of scans (same layout as Context.extract_lines)."""
    off = np.ascontiguousarray(point_offset, dtype=np.int64)
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    cnt = None if point_count is None else np.ascontiguousarray(point_count, dtype=np.int32)
    pz = None if point_z is None else np.ascontiguousarray(point_z, dtype=np.float64).reshape(-1)
    S = len(off) - 1 if cnt is None else len(cnt)
    n = np.zeros(S, np.int32)
    lines, abc, rng = np.zeros((S, max_lines, 4)), np.zeros((S, max_lines, 3)), np.zeros((S, max_lines, 2), np.int32)
    rc = lib().oracle_extract_lines(C.byref(lp), S, off.ctypes.data_as(abi.c_int64_p),
                                    cnt.ctypes.data_as(abi.c_int32_p) if cnt is not None else abi.c_int32_p(), _d(pts),
                                    _d(pz) if pz is not None else dp(), int(max_lines),
                                    n.ctypes.data_as(abi.c_int32_p), _d(lines), _d(abc), rng.ctypes.data_as(abi.c_int32_p))
    assert rc == 0
    return n, lines, abc, rng


def scan_to_points(ranges, headers, deskew=True):
    """convert::laser_to_point_times + sensor::laser::correct (same layout as Context.scan_to_points)."""
    rg = np.ascontiguousarray(ranges, dtype=np.float32)
    hd = np.ascontiguousarray(headers, dtype=abi.SCAN_HEADER_DTYPE)
    S, nb = rg.shape
    cnt = np.zeros(S, np.int32)
    pts, pz, pt = np.zeros((S, nb, 2)), np.zeros((S, nb)), np.zeros((S, nb))
    rc = lib().oracle_scan_to_points(S, nb, C.c_void_p(rg.ctypes.data), C.c_void_p(hd.ctypes.data), int(bool(deskew)),
                                     cnt.ctypes.data_as(abi.c_int32_p), _d(pts), _d(pz), _d(pt))
    assert rc == 0
    return cnt, pts, pz, pt


def match_lines(params, lp, n_lines1, lines1, n_lines2, lines2, pose1, pose2, kk=0, point_offset1=None, points1=None,
                index_range1=None, point_count1=None):
    """laser_manager::do_match for a batch of scan pairs (same layout as Context.match_lines)."""
    l1, l2 = np.ascontiguousarray(lines1, dtype=np.float64), np.ascontiguousarray(lines2, dtype=np.float64)
    n1, n2 = np.ascontiguousarray(n_lines1, dtype=np.int32), np.ascontiguousarray(n_lines2, dtype=np.int32)
    P, m1, m2 = len(n1), l1.shape[1], l2.shape[1]
    s1, s2 = _arr(pose1).reshape(P, 6), _arr(pose2).reshape(P, 6)
    i32 = abi.c_int32_p
    off = pts = rng = cnt = None
    if points1 is not None:
        off = np.ascontiguousarray(point_offset1, dtype=np.int64)
        pts = _arr(points1).reshape(-1, 2)
        rng = np.ascontiguousarray(index_range1, dtype=np.int32)
        cnt = None if point_count1 is None else np.ascontiguousarray(point_count1, dtype=np.int32)
    nm = np.zeros(P, np.int32)
    match = np.zeros((P, m2, 2), np.int32)
    rc = lib().oracle_match_lines(C.byref(params), C.byref(lp), P, int(kk), off.ctypes.data_as(abi.c_int64_p) if off is not None else abi.c_int64_p(),
                                  cnt.ctypes.data_as(i32) if cnt is not None else i32(), _d(pts) if pts is not None else dp(), m1,
                                  n1.ctypes.data_as(i32), _d(l1), rng.ctypes.data_as(i32) if rng is not None else i32(), m2,
                                  n2.ctypes.data_as(i32), _d(l2), _d(s1), _d(s2), nm.ctypes.data_as(i32), match.ctypes.data_as(i32))
    assert rc == 0
    return nm, match


def eval_edge_factor(tf12, weight, sqrt_info, pose_i, pose_j):
    """edge_factor (edge_factor.h:79-126): res[6], jac[6][12] over (p_i, q_i, p_j, q_j)."""
    res, jac = np.zeros(6), np.zeros((6, 12))
    lib().oracle_eval_edge_factor.argtypes = [dp, C.c_double, dp, dp, dp, dp, dp]
    lib().oracle_eval_edge_factor(_d(_arr(tf12).reshape(-1)), float(weight), _d(_arr(sqrt_info).reshape(-1)), _d(_arr(pose_i)), _d(_arr(pose_j)),
                                  _d(res), _d(jac))
    return res, jac


def pose_graph_solve(params, poses, edge_index, edge_tf, edge_weight, sqrt_info, ground_p=True, ground_q=True, fixed_pose=None):
    """keyframe_manager::solve (keyframe_manager.cpp:722-838): returns the optimised poses [K][6] and the summary.
    fixed_pose: the constant key frame (-1: none); None = index1 of the first edge (the reference's seq_edges[0])."""
    x = np.array(poses, dtype=np.float64).reshape(-1, 6).copy()
    ei = np.ascontiguousarray(edge_index, dtype=np.int32).reshape(-1, 2)
    et = _arr(edge_tf).reshape(-1, 12)
    ew = _arr(edge_weight).reshape(-1)
    summ = np.zeros(1, dtype=abi.SUMMARY_DTYPE)
    rc = lib().oracle_pose_graph_solve(C.byref(params), len(x), _d(x), len(ei), ei.ctypes.data_as(abi.c_int32_p), _d(et), _d(ew),
                                       _d(_arr(sqrt_info).reshape(-1)), int(ground_p), int(ground_q),
                                       (int(ei[0, 0]) if len(ei) else -1) if fixed_pose is None else int(fixed_pose), summ.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return x, summ


def fit_line(points):
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(3)
    lib().oracle_fit_line(_d(pts), len(pts), _d(out))
    return out


def max_threads():
    return int(lib().oracle_max_threads())


class OracleContext:
    """The oracle behind the method names of lvio2d_b200.solver.Context (set_windows / solve / get_states /
    marginalize), so that tests can run the `Solver` class flow on the CPU restatement and compare trajectories."""

    def __init__(self, params):
        self.params = params
        self.hb = None
        self.states = None

    def set_windows(self, hb):
        self.hb, self.states = hb, None

    def set_max_iterations(self, max_iters):
        self.params.max_iters = int(max_iters) if max_iters > 0 else 50

    def solve(self, want_summary=True):
        self.states, summ = solve(self.params, self.hb)
        return summ

    def get_states(self, out=None):
        return self.states if self.states is not None else np.array(self.hb["states"], dtype=np.float64)

    def marginalize(self):
        X0, J, r, _, _ = marginalize(self.params, self.hb, self.states)
        return X0, J, r

    # the laser front-end entry points, same signatures as lvio2d_b200.solver.Context
    def scan_to_points(self, ranges, headers, deskew=True, want_times=False):
        cnt, pts, pz, pt = scan_to_points(ranges, headers, deskew)
        return (cnt, pts, pz, pt) if want_times else (cnt, pts, pz)

    def extract_lines(self, line_params, point_offset, points, max_lines=256, point_count=None, point_z=None):
        return extract_lines(line_params, point_offset, points, max_lines, point_count, point_z)

    def match_lines(self, line_params, n_lines1, lines1, n_lines2, lines2, pose1, pose2, kk=0, point_offset1=None, points1=None,
                    index_range1=None, point_count1=None):
        return match_lines(self.params, line_params, n_lines1, lines1, n_lines2, lines2, pose1, pose2, kk, point_offset1, points1,
                           index_range1, point_count1)

    def close(self):
        pass
