"""ctypes wrapper of oracle/_ref/libref.so — TEST INFRASTRUCTURE.

libref.so is the REFERENCE'S OWN source text (src/factor/*.h, src/factor/solver.cpp, src/trajectory/laser_manager.cpp,
src/utilies/common.{h,cpp}, src/trajectory/sensor.h) compiled unmodified against the stub include tree
oracle/ref_stub (Eigen / Ceres / ROS are absent from this image) behind the C entry points of oracle/ref_driver.cpp.
It pins the oracle (and through it the CUDA path) to what the reference computes.

The reference caches its parameters in function-local singletons on first use (laser_noise, ground_noise, imu_noise,
wheel_noise, edge_noise), so ONE parameter set per process: corridor.yaml's values (lvio2d_b200.corridor_params).
`fast_mode` is read on every call and can be switched.

Built by `make -C oracle ref` where /root/reference exists; elsewhere (the GPU box) the prebuilt file is used.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import lvio2d_b200 as L  # noqa: E402
from lvio2d_b200 import abi  # noqa: E402

REF_SRC = "/root/reference/src"
_LIB_PATH = os.path.join(_ROOT, "oracle", "_ref", "libref.so")
_lib = None
dp = abi.c_double_p
ip = abi.c_int32_p
lp64 = abi.c_int64_p

# ref_set_params layout (oracle/ref_driver.cpp)
LOOP_SIGMA_P = (0.1, 0.1, 0.1)         # config/corridor.yaml:110
LOOP_SIGMA_Q = (0.01, 0.01, 0.01)      # :111
REF_MOTION_FILTER = (0.01, 0.01)       # :120-121
REF_N_ACCUMULATION = 100               # :122
LOOP_EDGE_K = 10.0                     # :106


def available():
    """True when libref.so exists or can be built here (the reference tree is present)."""
    return os.path.exists(_LIB_PATH) or os.path.exists(os.path.join(REF_SRC, "factor", "solver.cpp"))


def build():
    if os.path.exists(os.path.join(REF_SRC, "factor", "solver.cpp")):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s", "ref"])
    return _LIB_PATH


def _d(a):
    return a.ctypes.data_as(dp)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def param_vector(params, line_params, fast_mode=False, ref_n_accumulation=REF_N_ACCUMULATION):
    v = []
    v += list(params.T_imu_to_laser) + list(params.T_imu_to_wheel)
    v += [params.g, params.line_to_line_sigma, params.manifold_p_sigma, params.manifold_q_sigma]
    v += list(params.imu_noise_acc_sigma) + list(params.imu_bias_acc_sigma) + list(params.imu_noise_gyro_sigma) + list(params.imu_bias_gyro_sigma)
    v += list(params.wheel_sigma) + list(LOOP_SIGMA_P) + list(LOOP_SIGMA_Q)
    v += [1.0 if fast_mode else 0.0]
    v += [line_params.w_laser_each_scan, line_params.h_laser_each_scan, line_params.laser_resolution, line_params.line_continuous_threshold,
          line_params.line_max_tolerance_angle_deg, line_params.line_min_len, line_params.line_max_dis]
    v += [REF_MOTION_FILTER[0], REF_MOTION_FILTER[1], float(ref_n_accumulation), LOOP_EDGE_K]
    v += [0.0] * (70 - len(v))
    return np.array(v, dtype=np.float64)


def lib():
    """dlopen + bind + push corridor.yaml's parameters (once per process)."""
    global _lib
    if _lib is not None:
        return _lib
    build()
    lb = C.CDLL(_LIB_PATH)
    lb.ref_dis_from_line.restype = C.c_double
    lb.ref_dis_from_line.argtypes = [dp, dp, dp]
    lb.ref_laser_pair_weight.restype = C.c_double
    lb.ref_laser_pair_weight.argtypes = [dp, dp, dp, dp]
    lb.ref_eval_edge_factor.argtypes = [dp, C.c_double, dp, dp, dp, dp, dp]
    for name in ("ref_solver_new", "ref_scan_from_points", "ref_scan_from_lines", "ref_lm_new"):
        getattr(lb, name).restype = C.c_void_p
    lb.ref_scan_from_points.argtypes = [C.c_int, dp]
    lb.ref_scan_from_lines.argtypes = [C.c_int, dp]
    lb.ref_solver_free.argtypes = [C.c_void_p]
    lb.ref_scan_free.argtypes = [C.c_void_p]
    lb.ref_lm_free.argtypes = [C.c_void_p]
    lb.ref_solver_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, ip, lp64, dp, dp, dp, dp, dp, dp]
    lb.ref_solver_get_prior.argtypes = [C.c_void_p, dp, dp, dp]
    lb.ref_solver_set_prior.argtypes = [C.c_void_p, dp, dp, dp]
    lb.ref_solver_marg_system.argtypes = [C.c_void_p, C.c_int, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lb.ref_scan_num_lines.argtypes = [C.c_void_p]
    lb.ref_scan_num_corners.argtypes = [C.c_void_p]
    lb.ref_scan_get.argtypes = [C.c_void_p, dp, dp]
    lb.ref_do_match.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_int, ip]
    lb.ref_lm_add_scan.argtypes = [C.c_void_p, C.c_void_p, dp]
    lb.ref_lm_submap.argtypes = [C.c_void_p, C.c_int, dp, dp, C.POINTER(C.c_int)]
    lb.ref_lm_match_with_ref.argtypes = [C.c_void_p, C.c_void_p, dp, ip, dp]
    lb.ref_scan_to_points.argtypes = [C.c_int, C.POINTER(C.c_float), dp, C.c_int, dp, dp]
    lb.ref_imu_preintegrate.argtypes = [C.c_int, lp64, dp, dp, dp]
    lb.ref_imu_preintegrate_stamped.argtypes = [C.c_int, dp, dp, dp, dp]
    lb.ref_wheel_preintegrate.argtypes = [C.c_int, lp64, dp, dp]
    v = param_vector(L.corridor_params(), L.corridor_line_params())
    assert lb.ref_set_params(_d(v), len(v)) > 0
    _lib = lb
    return lb


def set_iteration_cap(cap):
    """Stub knob: cap every ceres::Solve of the reference text at `cap` LM iterations (0 = the reference's own setting)."""
    lib().ref_set_iteration_cap(int(cap))


def set_fast_mode(on):
    lib().ref_set_fast_mode(int(bool(on)))


# ---- primitives
def dis_from_line(p, p1, p2):
    return lib().ref_dis_from_line(_d(_arr(p)), _d(_arr(p1)), _d(_arr(p2)))


def exp_so3(v):
    R = np.zeros(9)
    lib().ref_exp_so3(_d(_arr(v)), _d(R))
    return R.reshape(3, 3)


def log_SO3(R):
    v = np.zeros(3)
    lib().ref_log_SO3(_d(_arr(R).ravel()), _d(v))
    return v


def normalize_so3(v):
    v = _arr(v).copy()
    lib().ref_normalize_so3(_d(v))
    return v


def so3_plus(theta, delta, want_jacobian=False):
    out, jac = np.zeros(3), np.zeros((3, 3))
    lib().ref_so3_plus(_d(_arr(theta)), _d(_arr(delta)), _d(out), _d(jac))
    return (out, jac) if want_jacobian else out


# ---- factors (residual, Jacobian over the stacked parameter blocks)
def eval_laser_factor(l1_p1, l1_p2, l2_p1, l2_p2, pose_i, pose_j):
    res, jac = np.zeros(2), np.zeros((2, 12))
    lib().ref_eval_laser_factor(_d(_arr(l1_p1)), _d(_arr(l1_p2)), _d(_arr(l2_p1)), _d(_arr(l2_p2)), _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def laser_pair_weight(l1_p1, l1_p2, l2_p1, l2_p2):
    return lib().ref_laser_pair_weight(_d(_arr(l1_p1)), _d(_arr(l1_p2)), _d(_arr(l2_p1)), _d(_arr(l2_p2)))


def eval_imu_factor(blob, state_i, state_j):
    res, jac = np.zeros(15), np.zeros((15, 30))
    lib().ref_eval_imu_factor(_d(_arr(blob)), _d(_arr(state_i)), _d(_arr(state_j)), _d(res), _d(jac))
    return res, jac


def eval_wheel_factor(blob, pose_i, pose_j):
    res, jac = np.zeros(3), np.zeros((3, 12))
    lib().ref_eval_wheel_factor(_d(_arr(blob)), _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac))
    return res, jac


def eval_ground_factors(pose):
    res, jac = np.zeros(2), np.zeros((2, 6))
    lib().ref_eval_ground_factors(_d(_arr(pose)), _d(res), _d(jac))
    return res, jac


def eval_prior_factor(X0, J, state):
    res, jac = np.zeros(15), np.zeros((15, 15))
    lib().ref_eval_prior_factor(_d(_arr(X0)), _d(_arr(J).ravel()), _d(_arr(state)), _d(res), _d(jac))
    return res, jac


def eval_edge_factor(tf12, weight, pose_i, pose_j):
    """-> res[6], jac[6][12], edge_noise::J[6][6] as the reference builds it"""
    res, jac, Jn = np.zeros(6), np.zeros((6, 12)), np.zeros((6, 6))
    lib().ref_eval_edge_factor(_d(_arr(tf12).ravel()), float(weight), _d(_arr(pose_i)), _d(_arr(pose_j)), _d(res), _d(jac), _d(Jn))
    return res, jac, Jn


# ---- preintegration
def imu_preintegrate(sample_offset, samples, bias0):
    off = np.ascontiguousarray(sample_offset, dtype=np.int64)
    n = off.size - 1
    out = np.zeros((n, abi.IMU_BLOB))
    lib().ref_imu_preintegrate(n, off.ctypes.data_as(lp64), _d(_arr(samples)), _d(_arr(bias0)), _d(out))
    return out


def imu_preintegrate_stamped(stamps, acc_gyro, bias):
    out = np.zeros(abi.IMU_BLOB)
    ag = _arr(acc_gyro)
    lib().ref_imu_preintegrate_stamped(len(ag), _d(_arr(stamps)), _d(ag), _d(_arr(bias)), _d(out))
    return out


def wheel_preintegrate(step_offset, steps):
    off = np.ascontiguousarray(step_offset, dtype=np.int64)
    n = off.size - 1
    out = np.zeros((n, abi.WHEEL_BLOB))
    lib().ref_wheel_preintegrate(n, off.ctypes.data_as(lp64), _d(_arr(steps)), _d(out))
    return out


# ---- lvio_2d::solver on FrameInfo lists (the same objects lvio2d_b200.solver.Solver takes)
class RefSolver:
    """The reference's solver object (src/factor/solver.cpp) behind the interface of lvio2d_b200.solver.Solver."""

    def __init__(self, fast_mode=False):
        self.fast_mode = bool(fast_mode)
        self._h = C.c_void_p(lib().ref_solver_new())
        self.last_summary = None

    def close(self):
        if self._h:
            lib().ref_solver_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, which, frames):
        set_fast_mode(self.fast_mode)
        n = len(frames)
        states = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in frames]).astype(np.float64)
        has = np.zeros(n, np.int32)
        off = np.zeros(n + 1, np.int64)
        pairs, mpose = [], np.zeros((n, 12))
        for i, f in enumerate(frames):
            m = f.laser_match
            if m is not None:
                has[i] = 1
                for l1, l2 in zip(m.lines1, m.lines2):
                    pairs.append(np.concatenate([l1.p1, l1.p2, l2.p1, l2.p2]))
                mpose[i] = np.concatenate([m.p1, m.q1, m.p2, m.q2])
            off[i + 1] = len(pairs)
        pairs = _arr(np.array(pairs).reshape(-1, 12)) if pairs else np.zeros((1, 12))
        imu = np.zeros((n, abi.IMU_BLOB))
        wheel = np.zeros((n, abi.WHEEL_BLOB))
        for i in range(1, n):
            imu[i], wheel[i] = frames[i].imu_observation_result, frames[i].wheel_observation_result
        sqrtH, summ = np.zeros((6, 6)), np.zeros(8)
        rc = lib().ref_solver_run(self._h, which, n, _d(states), has.ctypes.data_as(ip), off.ctypes.data_as(lp64), _d(pairs), _d(mpose), _d(imu),
                                  _d(wheel), _d(sqrtH), _d(summ))
        assert rc == 0
        for i, f in enumerate(frames):
            s = states[i]
            f.p[:], f.q[:], f.v[:], f.bs[:] = s[0:3], s[3:6], s[6:9], s[9:15]
            if f.laser_match is not None:
                m = f.laser_match
                m.p1, m.q1, m.p2, m.q2 = mpose[i, 0:3].copy(), mpose[i, 3:6].copy(), mpose[i, 6:9].copy(), mpose[i, 9:12].copy()
        # same record layout as lvio2d_b200.solver.Context.solve() returns; Ceres reports costs including the fixed cost
        # of residual blocks without a variable parameter block, the product reports the reduced program's cost
        out = np.zeros(1, dtype=abi.SUMMARY_DTYPE)
        out["iterations"], out["termination"] = int(summ[0]), int(summ[1])
        out["num_successful_steps"], out["num_unsuccessful_steps"] = int(summ[2]), int(summ[3])
        out["initial_cost"], out["final_cost"], out["final_radius"] = summ[4] - summ[7], summ[5] - summ[7], summ[6]
        self.last_summary = out
        self.fixed_cost = float(summ[7])
        return sqrtH

    def solve(self, frame_infos, feature_infos=None):
        self._run(0, frame_infos)

    def init_solve(self, frame_infos, feature_infos=None):
        self._run(1, frame_infos)

    def marginalization(self, frame_infos, feature_infos=None):
        sqrtH = self._run(2, frame_infos)
        if not self.fast_mode:
            frame_infos[-1].sqrt_H = sqrtH

    # the attributes lvio2d_b200.solver.Solver keeps (replay.run_tracking_lockstep copies them between solvers)
    @property
    def has_linearized_block(self):
        return self.prior is not None

    @property
    def linearized_X(self):
        return None if self.prior is None else self.prior[0]

    @property
    def linearized_jacobians(self):
        return None if self.prior is None else self.prior[1]

    @property
    def linearized_residuals(self):
        return None if self.prior is None else self.prior[2]

    @property
    def prior(self):
        """(X0, J, r) kept by solver::marginalization, or None"""
        X0, J, r = np.zeros(15), np.zeros((15, 15)), np.zeros(15)
        if not lib().ref_solver_get_prior(self._h, _d(X0), _d(J), _d(r)):
            return None
        return X0, J, r

    def set_prior(self, X0, J, r=None):
        r = np.zeros(15) if r is None else r
        lib().ref_solver_set_prior(self._h, _d(_arr(X0)), _d(_arr(J).ravel()), _d(_arr(r)))

    def marg_system(self):
        """marginalization_matrix (solver.cpp:4-40) of the J, R the last marginalization() assembled -> (Delta_H, Delta_g, J shape)"""
        dH, dg = np.zeros((15, 15)), np.zeros(15)
        rows, cols = C.c_int(), C.c_int()
        rc = lib().ref_solver_marg_system(self._h, 15, _d(dH), _d(dg), C.byref(rows), C.byref(cols))
        assert rc == 0
        return dH, dg, (rows.value, cols.value)


# ---- laser front-end
def scan_to_points(ranges, header, deskew=True):
    """convert::laser_to_point_times (+ sensor::laser::correct).  header: one row of abi.SCAN_HEADER_DTYPE; range_min/max are
    not read by the reference's conversion.  -> points [n][3], times [n]"""
    r = np.ascontiguousarray(ranges, dtype=np.float32)
    h = np.array([header["angle_min"], header["angle_increment"], header["time_increment"], 0.0, 0.0, header["stamp"]] + list(header["linear"]) +
                 list(header["angular"]), dtype=np.float64)
    pts, times = np.zeros((len(r), 3)), np.zeros(len(r))
    n = lib().ref_scan_to_points(len(r), r.ctypes.data_as(C.POINTER(C.c_float)), _d(h), int(bool(deskew)), _d(pts), _d(times))
    return pts[:n].copy(), times[:n].copy()


class RefScan:
    """lvio_2d::scan built by laser_manager::spawn_scan (from points) or by scan::add_line(p1, p2, false) (from lines)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    @classmethod
    def from_points(cls, points3):
        p = _arr(points3).reshape(-1, 3)
        return cls(lib().ref_scan_from_points(len(p), _d(p)))

    @classmethod
    def from_lines(cls, lines6):
        l6 = _arr(lines6).reshape(-1, 6)
        return cls(lib().ref_scan_from_lines(len(l6), _d(l6)))

    def __del__(self):
        try:
            if self._h:
                lib().ref_scan_free(self._h)
                self._h = None
        except Exception:
            pass

    def num_lines(self):
        return int(lib().ref_scan_num_lines(self._h))

    def lines(self):
        """-> [n][9] = p1 p2 abc, corners [m][3]"""
        n, m = lib().ref_scan_num_lines(self._h), lib().ref_scan_num_corners(self._h)
        out, cor = np.zeros((max(n, 1), 9)), np.zeros((max(m, 1), 3))
        lib().ref_scan_get(self._h, _d(out), _d(cor))
        return out[:n], cor[:m]


def do_match(scan1, scan2, pose1, pose2, kk=0):
    n2 = lib().ref_scan_num_lines(scan2._h)
    pairs = np.zeros((max(n2, 1), 2), np.int32)
    n = lib().ref_do_match(scan1._h, scan2._h, _d(_arr(pose1)), _d(_arr(pose2)), int(kk), pairs.ctypes.data_as(ip))
    return pairs[:n].copy()


class RefLaserManager:
    """lvio_2d::laser_manager: add_scan / match_with_ref and a view of its two reference sub-maps."""

    def __init__(self):
        self._h = C.c_void_p(lib().ref_lm_new())
        self._scans = []

    def __del__(self):
        try:
            if self._h:
                lib().ref_lm_free(self._h)
                self._h = None
        except Exception:
            pass

    def add_scan(self, scan, pose):
        self._scans.append(scan)
        lib().ref_lm_add_scan(self._h, scan._h, _d(_arr(pose)))

    def submap(self, which=0):
        """-> None or (pose[6], lines [n][6], current_count)"""
        cnt = C.c_int()
        n = lib().ref_lm_submap(self._h, which, None, None, C.byref(cnt))
        if n < 0:
            return None
        pose, lines = np.zeros(6), np.zeros((max(n, 1), 6))
        lib().ref_lm_submap(self._h, which, _d(pose), _d(lines), C.byref(cnt))
        return pose, lines[:n], cnt.value

    def match_with_ref(self, scan, pose):
        n2 = lib().ref_scan_num_lines(scan._h)
        pairs, ref_pose = np.zeros((max(n2, 1), 2), np.int32), np.zeros(6)
        n = lib().ref_lm_match_with_ref(self._h, scan._h, _d(_arr(pose)), pairs.ctypes.data_as(ip), _d(ref_pose))
        return pairs[:n].copy(), ref_pose
