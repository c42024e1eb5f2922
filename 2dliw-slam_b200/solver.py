"""Host-side mirror of the reference's solver interface, on top of the C ABI (include/lvio2d.h).

Two layers:
  * `Context` — a thin ctypes binding of liblvio2d.so, one method per C entry point.  This is what a
    reference maintainer's C++ shim calls (shim/lvio2d_solver_shim.h); Python only plays the role of the
    host language here because the reference's own host toolchain (ROS/catkin, Eigen) is absent.
  * `Solver` / `FrameInfo` / `LaserMatch` — the reference's `lvio_2d::solver` surface
    (reference src/factor/solver.h:71-79: solve, init_solve, marginalization on a deque of frame_info,
    src/trajectory/trajectory_type.h:9-75) with the same names, argument meaning and in-place write-back.

The compute path is CUDA only.  Importing this module never builds or falls back to anything: if
csrc/liblvio2d.so is missing or no sm_100 device is present, `Context()` raises.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# LVIO2D_LIB: another build of the same CUDA library (kernel-variant experiments, scripts/variants.sh)
LIB_PATH = os.environ.get("LVIO2D_LIB") or os.path.join(_HERE, "csrc", "liblvio2d.so")
_lib = None

EXPORTS = [
    "lvio2d_create", "lvio2d_destroy", "lvio2d_strerror", "lvio2d_last_error", "lvio2d_stream", "lvio2d_set_windows",
    "lvio2d_bind_windows", "lvio2d_reset_states", "lvio2d_solve", "lvio2d_solve_async", "lvio2d_sync", "lvio2d_get_summaries",
    "lvio2d_get_states", "lvio2d_set_point_shard", "lvio2d_solve_begin", "lvio2d_eval_laser", "lvio2d_reduce_buffer",
    "lvio2d_set_reduce_buffer", "lvio2d_lm_step", "lvio2d_linearize", "lvio2d_marginalize", "lvio2d_imu_preintegrate",
    "lvio2d_wheel_preintegrate", "lvio2d_eval_laser_factor", "lvio2d_eval_imu_factor", "lvio2d_eval_wheel_factor",
    "lvio2d_eval_ground_factors", "lvio2d_set_profiling", "lvio2d_get_profile", "lvio2d_set_windows_async", "lvio2d_get_states_async",
    "lvio2d_extract_lines", "lvio2d_scan_to_points", "lvio2d_match_lines", "lvio2d_pose_graph_solve", "lvio2d_eval_edge_factor",
    "lvio2d_measure_fp64_peak", "lvio2d_set_max_iterations", "lvio2d_set_windows_wire", "lvio2d_submap_create", "lvio2d_submap_destroy",
    "lvio2d_submap_reset", "lvio2d_submap_add_scan", "lvio2d_submap_get", "lvio2d_submap_device", "lvio2d_submap_match",
]


class Lvio2dError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: status {status}{' (' + detail + ')' if detail else ''}")


def load_library(path=LIB_PATH):
    """dlopen the CUDA library of the C ABI.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc, sm_100a). There is no CPU fallback for this path.")
    lib = C.CDLL(path)
    vp, dp = C.c_void_p, abi.c_double_p
    lib.lvio2d_create.argtypes = [C.POINTER(vp), C.POINTER(abi.Params)]
    lib.lvio2d_destroy.argtypes = [vp]
    lib.lvio2d_destroy.restype = None
    lib.lvio2d_strerror.argtypes = [C.c_int]
    lib.lvio2d_strerror.restype = C.c_char_p
    lib.lvio2d_last_error.argtypes = [vp]
    lib.lvio2d_last_error.restype = C.c_char_p
    lib.lvio2d_stream.argtypes = [vp]
    lib.lvio2d_stream.restype = vp
    lib.lvio2d_set_windows.argtypes = [vp, C.POINTER(abi.WindowBatch)]
    lib.lvio2d_bind_windows.argtypes = [vp, C.POINTER(abi.WindowBatch)]
    lib.lvio2d_reset_states.argtypes = [vp, vp]
    lib.lvio2d_solve.argtypes = [vp, vp]
    lib.lvio2d_solve_async.argtypes = [vp]
    lib.lvio2d_sync.argtypes = [vp]
    lib.lvio2d_get_summaries.argtypes = [vp, vp]
    lib.lvio2d_get_states.argtypes = [vp, vp]
    lib.lvio2d_set_point_shard.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lvio2d_solve_begin.argtypes = [vp]
    lib.lvio2d_eval_laser.argtypes = [vp]
    lib.lvio2d_reduce_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    lib.lvio2d_set_reduce_buffer.argtypes = [vp, vp, C.c_int64]
    lib.lvio2d_lm_step.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.lvio2d_linearize.argtypes = [vp, C.c_int32, dp, dp, dp]
    lib.lvio2d_marginalize.argtypes = [vp, dp, dp, dp]
    lib.lvio2d_imu_preintegrate.argtypes = [vp, C.c_int32, abi.c_int64_p, dp, dp, dp]
    lib.lvio2d_wheel_preintegrate.argtypes = [vp, C.c_int32, abi.c_int64_p, dp, dp]
    lib.lvio2d_eval_laser_factor.argtypes = [vp] + [dp] * 8
    lib.lvio2d_eval_imu_factor.argtypes = [vp] + [dp] * 5
    lib.lvio2d_eval_wheel_factor.argtypes = [vp] + [dp] * 5
    lib.lvio2d_eval_ground_factors.argtypes = [vp] + [dp] * 3
    lib.lvio2d_set_profiling.argtypes = [vp, C.c_int32]
    lib.lvio2d_set_windows_async.argtypes = [vp, C.POINTER(abi.WindowBatch)]
    lib.lvio2d_get_states_async.argtypes = [vp, vp]
    lib.lvio2d_extract_lines.argtypes = [vp, C.POINTER(abi.LineParams), C.c_int32, abi.c_int64_p, abi.c_int32_p, dp, dp, C.c_int32,
                                         abi.c_int32_p, dp, dp, abi.c_int32_p, C.c_int32]
    lib.lvio2d_match_lines.argtypes = [vp, C.POINTER(abi.LineParams), C.c_int32, C.c_int32, abi.c_int64_p, abi.c_int32_p, dp, C.c_int32,
                                       abi.c_int32_p, dp, abi.c_int32_p, C.c_int32, abi.c_int32_p, dp, dp, dp, abi.c_int32_p, abi.c_int32_p,
                                       C.c_int32]
    lib.lvio2d_scan_to_points.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, abi.c_int32_p, dp, dp, dp, C.c_int32]
    lib.lvio2d_submap_create.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(abi.LineParams), C.c_double, C.c_double, C.c_int32, C.POINTER(vp)]
    lib.lvio2d_submap_destroy.argtypes = [vp]
    lib.lvio2d_submap_destroy.restype = None
    lib.lvio2d_submap_reset.argtypes = [vp]
    lib.lvio2d_submap_add_scan.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32]
    lib.lvio2d_submap_get.argtypes = [vp, C.c_int32, abi.c_int32_p, dp, abi.c_int32_p, dp]
    lib.lvio2d_submap_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.lvio2d_submap_match.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32]
    lib.lvio2d_get_profile.argtypes = [vp, dp]
    lib.lvio2d_measure_fp64_peak.argtypes = [vp, dp]
    lib.lvio2d_set_max_iterations.argtypes = [vp, C.c_int32]
    lib.lvio2d_set_windows_wire.argtypes = [vp, C.POINTER(abi.WindowBatch), C.POINTER(abi.ScanWireStruct), C.c_int32]
    lib.lvio2d_pose_graph_solve.argtypes = [vp, C.c_int32, dp, C.c_int32, abi.c_int32_p, dp, dp, dp, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.lvio2d_eval_edge_factor.argtypes = [vp, dp, C.c_double, dp, dp, dp, dp, dp]
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(abi.c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """lvio2d_ctx: one solver instance = one CUDA stream + device buffers (include/lvio2d.h)."""

    def __init__(self, params):
        self.lib = load_library()
        self.params = params
        self._h = C.c_void_p()
        rc = self.lib.lvio2d_create(C.byref(self._h), C.byref(params))
        if rc != abi.OK:
            self._h = C.c_void_p()
            raise Lvio2dError(rc, "lvio2d_create", self.lib.lvio2d_strerror(rc).decode())
        self.n_windows = self.n_frames = 0
        self._keep = None
        self._submaps = []

    def close(self):
        if self._h:
            for ref in self._submaps:      # sub-maps enqueue on this context's stream: they go first
                sm = ref()
                if sm is not None:
                    sm.close()
            self._submaps = []
            self.lib.lvio2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != abi.OK:
            raise Lvio2dError(rc, what, self.lib.lvio2d_last_error(self._h).decode() or self.lib.lvio2d_strerror(rc).decode())

    @property
    def stream(self):
        """cudaStream_t (as int) all work of this context is enqueued on."""
        return int(self.lib.lvio2d_stream(self._h) or 0)

    # ---- problem upload
    def set_windows(self, host_batch):
        s = host_batch.struct()
        self._check(self.lib.lvio2d_set_windows(self._h, C.byref(s)), "lvio2d_set_windows")
        self.n_windows, self.n_frames = host_batch.n_windows, host_batch.n_frames

    def set_windows_async(self, host_batch):
        """Enqueue the upload only; `host_batch` (pinned) must stay alive and unchanged until sync()."""
        s = host_batch.struct()
        self._check(self.lib.lvio2d_set_windows_async(self._h, C.byref(s)), "lvio2d_set_windows_async")
        self.n_windows, self.n_frames = host_batch.n_windows, host_batch.n_frames
        self._keep = host_batch

    def set_windows_wire(self, host_batch, wire, async_=False):
        """lvio2d_set_windows_wire: `host_batch` without points / point_line / point_offset, the laser input as abi.ScanWire."""
        s, ws = host_batch.struct(), wire.struct()
        self._check(self.lib.lvio2d_set_windows_wire(self._h, C.byref(s), C.byref(ws), int(bool(async_))), "lvio2d_set_windows_wire")
        self.n_windows, self.n_frames = host_batch.n_windows, host_batch.n_frames
        self._keep = (host_batch, wire)

    def get_states_async(self, out):
        self._check(self.lib.lvio2d_get_states_async(self._h, out.ctypes.data), "lvio2d_get_states_async")

    def bind_windows(self, device_struct, keepalive=None):
        """`device_struct`: abi.WindowBatch whose pointers are device pointers; `keepalive` owns the memory."""
        self._check(self.lib.lvio2d_bind_windows(self._h, C.byref(device_struct)), "lvio2d_bind_windows")
        self.n_windows, self.n_frames = device_struct.n_windows, device_struct.n_frames
        self._keep = keepalive

    def reset_states(self, states):
        st = _f64(states)
        assert st.size == self.n_windows * self.n_frames * 15
        self._check(self.lib.lvio2d_reset_states(self._h, st.ctypes.data), "lvio2d_reset_states")

    # ---- solve
    def solve(self, want_summary=True):
        if want_summary:
            summ = np.zeros(self.n_windows, dtype=abi.SUMMARY_DTYPE)
            self._check(self.lib.lvio2d_solve(self._h, summ.ctypes.data), "lvio2d_solve")
            return summ
        self._check(self.lib.lvio2d_solve(self._h, None), "lvio2d_solve")
        return None

    def solve_async(self):
        self._check(self.lib.lvio2d_solve_async(self._h), "lvio2d_solve_async")

    def sync(self):
        self._check(self.lib.lvio2d_sync(self._h), "lvio2d_sync")

    def get_summaries(self):
        summ = np.zeros(self.n_windows, dtype=abi.SUMMARY_DTYPE)
        self._check(self.lib.lvio2d_get_summaries(self._h, summ.ctypes.data), "lvio2d_get_summaries")
        return summ

    def get_states(self, out=None):
        if out is None:
            out = np.zeros((self.n_windows * self.n_frames, 15))
        self._check(self.lib.lvio2d_get_states(self._h, out.ctypes.data), "lvio2d_get_states")
        return out

    # ---- split phase (multi-GPU)
    def set_point_shard(self, rank, world):
        self._check(self.lib.lvio2d_set_point_shard(self._h, rank, world), "lvio2d_set_point_shard")

    def solve_begin(self):
        self._check(self.lib.lvio2d_solve_begin(self._h), "lvio2d_solve_begin")

    def eval_laser(self):
        self._check(self.lib.lvio2d_eval_laser(self._h), "lvio2d_eval_laser")

    def reduce_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.lvio2d_reduce_buffer(self._h, C.byref(p), C.byref(n)), "lvio2d_reduce_buffer")
        return int(p.value or 0), int(n.value)

    def set_reduce_buffer(self, device_ptr, count):
        self._check(self.lib.lvio2d_set_reduce_buffer(self._h, device_ptr, count), "lvio2d_set_reduce_buffer")

    def lm_step(self, want_active=False):
        if want_active:
            act = C.c_int32()
            self._check(self.lib.lvio2d_lm_step(self._h, C.byref(act)), "lvio2d_lm_step")
            return int(act.value)
        self._check(self.lib.lvio2d_lm_step(self._h, None), "lvio2d_lm_step")
        return None

    def set_max_iterations(self, max_iters):
        """max_num_iterations of the following solves (solver.cpp:800-801 vs :161-168)."""
        self._check(self.lib.lvio2d_set_max_iterations(self._h, int(max_iters)), "lvio2d_set_max_iterations")
        self.params.max_iters = int(max_iters) if max_iters > 0 else 50

    # ---- measurement
    def set_profiling(self, on=True):
        self._check(self.lib.lvio2d_set_profiling(self._h, int(on)), "lvio2d_set_profiling")

    def get_profile(self):
        out = np.zeros(8)
        self._check(self.lib.lvio2d_get_profile(self._h, _d(out)), "lvio2d_get_profile")
        return dict(scan_ms=out[0], scan_launches=int(out[1]), window_ms=out[2], window_launches=int(out[3]),
                    kernel_launches=int(out[4]), scan_bytes_per_launch=out[5], factor_ms=out[6], factor_launches=int(out[7]))

    def measure_fp64_peak(self):
        """Measured fp64 roofline denominators of this device: vector-pipe DFMA and tensor-pipe DMMA TFLOP/s."""
        out = np.zeros(4)
        self._check(self.lib.lvio2d_measure_fp64_peak(self._h, _d(out)), "lvio2d_measure_fp64_peak")
        return dict(dfma_tflops=out[0], dmma_tflops=out[1], dfma_ms=out[2], dmma_ms=out[3])

    # ---- linearisation / marginalisation
    def linearize(self, mode=0):
        B, dim = self.n_windows, 15 * self.n_frames
        H, g, cost = np.zeros((B, dim, dim)), np.zeros((B, dim)), np.zeros(B)
        self._check(self.lib.lvio2d_linearize(self._h, mode, _d(H), _d(g), _d(cost)), "lvio2d_linearize")
        return H, g, cost

    def marginalize(self):
        B = self.n_windows
        X0, J, r = np.zeros((B, 15)), np.zeros((B, 15, 15)), np.zeros((B, 15))
        self._check(self.lib.lvio2d_marginalize(self._h, _d(X0), _d(J), _d(r)), "lvio2d_marginalize")
        return X0, J, r

    # ---- preintegration
    def imu_preintegrate(self, sample_offset, samples, bias0):
        off = np.ascontiguousarray(sample_offset, dtype=np.int64)
        n = off.size - 1
        out = np.zeros((n, abi.IMU_BLOB))
        sm, b0 = _f64(samples), _f64(bias0)
        self._check(self.lib.lvio2d_imu_preintegrate(self._h, n, off.ctypes.data_as(abi.c_int64_p), _d(sm), _d(b0), _d(out)),
                    "lvio2d_imu_preintegrate")
        return out

    def wheel_preintegrate(self, step_offset, steps):
        off = np.ascontiguousarray(step_offset, dtype=np.int64)
        n = off.size - 1
        out = np.zeros((n, abi.WHEEL_BLOB))
        st = _f64(steps)
        self._check(self.lib.lvio2d_wheel_preintegrate(self._h, n, off.ctypes.data_as(abi.c_int64_p), _d(st), _d(out)),
                    "lvio2d_wheel_preintegrate")
        return out

    def preintegrate_batch(self, sensor_batch):
        """synth.SensorBatch -> abi.HostBatch through the device preintegrators."""
        if sensor_batch.n_frames > 1:
            imu = self.imu_preintegrate(sensor_batch.imu_offset, sensor_batch.imu_samples, sensor_batch.bias0)
            wheel = self.wheel_preintegrate(sensor_batch.wheel_offset, sensor_batch.wheel_steps)
        else:
            imu = wheel = None
        return sensor_batch.host_batch(imu, wheel)

    # ---- per-factor hooks (auto_diff::compute_res_and_jacobi, reference src/utilies/common.h:201-217)
    def extract_lines(self, line_params, point_offset, points, max_lines=256, point_count=None, point_z=None):
        """laser_manager::spawn_scan for a batch of scans (host buffers): returns n_lines [S], lines [S][max_lines][4],
        abc [S][max_lines][3], index_range [S][max_lines][2].  With `point_count`, scan s owns point_count[s] points
        from point_offset[s] (the fixed-stride layout scan_to_points writes)."""
        off = np.ascontiguousarray(point_offset, dtype=np.int64)
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        cnt = None if point_count is None else np.ascontiguousarray(point_count, dtype=np.int32)
        pz = None if point_z is None else np.ascontiguousarray(point_z, dtype=np.float64).reshape(-1)
        S = len(off) - 1 if cnt is None else len(cnt)
        n = np.zeros(S, np.int32)
        lines, abc = np.zeros((S, max_lines, 4)), np.zeros((S, max_lines, 3))
        rng = np.zeros((S, max_lines, 2), np.int32)
        self._check(self.lib.lvio2d_extract_lines(
            self._h, C.byref(line_params), S, off.ctypes.data_as(abi.c_int64_p),
            cnt.ctypes.data_as(abi.c_int32_p) if cnt is not None else abi.c_int32_p(), _d(pts),
            _d(pz) if pz is not None else abi.c_double_p(), int(max_lines), n.ctypes.data_as(abi.c_int32_p), _d(lines), _d(abc),
            rng.ctypes.data_as(abi.c_int32_p), 0), "lvio2d_extract_lines")
        return n, lines, abc, rng

    def extract_lines_device(self, line_params, n_scans, point_offset_ptr, points_ptr, max_lines, n_lines_ptr, lines_ptr, abc_ptr,
                             range_ptr, point_count_ptr=0, point_z_ptr=0):
        """Same with device pointers (integers); enqueued on the context's stream."""
        vp = C.c_void_p
        self._check(self.lib.lvio2d_extract_lines(
            self._h, C.byref(line_params), int(n_scans), C.cast(vp(point_offset_ptr), abi.c_int64_p),
            C.cast(vp(point_count_ptr or None), abi.c_int32_p), C.cast(vp(points_ptr), abi.c_double_p),
            C.cast(vp(point_z_ptr or None), abi.c_double_p), int(max_lines), C.cast(vp(n_lines_ptr), abi.c_int32_p),
            C.cast(vp(lines_ptr), abi.c_double_p), C.cast(vp(abc_ptr), abi.c_double_p), C.cast(vp(range_ptr), abi.c_int32_p), 1),
            "lvio2d_extract_lines")

    def match_lines(self, line_params, n_lines1, lines1, n_lines2, lines2, pose1, pose2, kk=0, point_offset1=None, points1=None,
                    index_range1=None, point_count1=None):
        """laser_manager::do_match for a batch of scan pairs (host buffers).  lines* [P][max_lines*][4] as returned by
        extract_lines; with points1 / index_range1 scan 1's cells are those of its lines' own points, otherwise the
        0.05 m samples along each segment.  Returns n_match [P], match [P][max_lines2][2] = (line of scan 1, line of scan 2)."""
        l1, l2 = np.ascontiguousarray(lines1, dtype=np.float64), np.ascontiguousarray(lines2, dtype=np.float64)
        n1, n2 = np.ascontiguousarray(n_lines1, dtype=np.int32), np.ascontiguousarray(n_lines2, dtype=np.int32)
        P, m1, m2 = len(n1), l1.shape[1], l2.shape[1]
        s1, s2 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(P, 6), np.ascontiguousarray(pose2, dtype=np.float64).reshape(P, 6)
        i32 = abi.c_int32_p
        off = pts = rng = cnt = None
        if points1 is not None:
            off = np.ascontiguousarray(point_offset1, dtype=np.int64)
            pts = np.ascontiguousarray(points1, dtype=np.float64).reshape(-1, 2)
            rng = np.ascontiguousarray(index_range1, dtype=np.int32)
            cnt = None if point_count1 is None else np.ascontiguousarray(point_count1, dtype=np.int32)
        nm = np.zeros(P, np.int32)
        match = np.zeros((P, m2, 2), np.int32)
        self._check(self.lib.lvio2d_match_lines(
            self._h, C.byref(line_params), P, int(kk), off.ctypes.data_as(abi.c_int64_p) if off is not None else abi.c_int64_p(),
            cnt.ctypes.data_as(i32) if cnt is not None else i32(), _d(pts) if pts is not None else abi.c_double_p(), m1,
            n1.ctypes.data_as(i32), _d(l1), rng.ctypes.data_as(i32) if rng is not None else i32(), m2, n2.ctypes.data_as(i32), _d(l2),
            _d(s1), _d(s2), nm.ctypes.data_as(i32), match.ctypes.data_as(i32), 0), "lvio2d_match_lines")
        return nm, match

    def submap(self, line_params, n_managers=1, line_cap=16384, ref_motion_filter_p=0.01, ref_motion_filter_q=0.01, ref_n_accumulation=100):
        """A batch of device-resident reference sub-maps (laser_manager::add_scan / match_with_ref), see Submap."""
        sm = Submap(self, line_params, n_managers, line_cap, ref_motion_filter_p, ref_motion_filter_q, ref_n_accumulation)
        self._submaps.append(weakref.ref(sm))
        return sm

    def scan_to_points(self, ranges, headers, deskew=True, want_times=False):
        """convert::laser_to_point_times + sensor::laser::correct for a batch of scans (host buffers).  ranges [S][n_beams]
        float32, headers: numpy array of abi.SCAN_HEADER_DTYPE.  Returns point_count [S], points [S][n_beams][2],
        point_z [S][n_beams] (and point_time)."""
        rg = np.ascontiguousarray(ranges, dtype=np.float32)
        hd = np.ascontiguousarray(headers, dtype=abi.SCAN_HEADER_DTYPE)
        S, nb = rg.shape
        cnt = np.zeros(S, np.int32)
        pts, pz = np.zeros((S, nb, 2)), np.zeros((S, nb))
        pt = np.zeros((S, nb)) if want_times else None
        self._check(self.lib.lvio2d_scan_to_points(self._h, S, nb, rg.ctypes.data, hd.ctypes.data, int(bool(deskew)),
                                                   cnt.ctypes.data_as(abi.c_int32_p), _d(pts), _d(pz),
                                                   _d(pt) if pt is not None else abi.c_double_p(), 0), "lvio2d_scan_to_points")
        return (cnt, pts, pz, pt) if want_times else (cnt, pts, pz)

    def scan_to_points_device(self, n_scans, n_beams, ranges_ptr, headers_ptr, deskew, count_ptr, points_ptr, z_ptr, time_ptr=0):
        vp = C.c_void_p
        self._check(self.lib.lvio2d_scan_to_points(self._h, int(n_scans), int(n_beams), ranges_ptr, headers_ptr, int(bool(deskew)),
                                                   C.cast(vp(count_ptr), abi.c_int32_p), C.cast(vp(points_ptr), abi.c_double_p),
                                                   C.cast(vp(z_ptr), abi.c_double_p), C.cast(vp(time_ptr or None), abi.c_double_p), 1),
                    "lvio2d_scan_to_points")

    def eval_laser_factor(self, l1_p1, l1_p2, l2_p1, l2_p2, pose_i, pose_j):
        res, jac = np.zeros(2), np.zeros((2, 12))
        a = [_f64(x) for x in (l1_p1, l1_p2, l2_p1, l2_p2, pose_i, pose_j)]
        self._check(self.lib.lvio2d_eval_laser_factor(self._h, *[_d(x) for x in a], _d(res), _d(jac)), "lvio2d_eval_laser_factor")
        return res, jac

    def eval_imu_factor(self, blob, state_i, state_j):
        res, jac = np.zeros(15), np.zeros((15, 30))
        a = [_f64(x) for x in (blob, state_i, state_j)]
        self._check(self.lib.lvio2d_eval_imu_factor(self._h, *[_d(x) for x in a], _d(res), _d(jac)), "lvio2d_eval_imu_factor")
        return res, jac

    def eval_wheel_factor(self, blob, pose_i, pose_j):
        res, jac = np.zeros(3), np.zeros((3, 12))
        a = [_f64(x) for x in (blob, pose_i, pose_j)]
        self._check(self.lib.lvio2d_eval_wheel_factor(self._h, *[_d(x) for x in a], _d(res), _d(jac)), "lvio2d_eval_wheel_factor")
        return res, jac

    def eval_ground_factors(self, pose):
        res, jac = np.zeros(2), np.zeros((2, 6))
        p = _f64(pose)
        self._check(self.lib.lvio2d_eval_ground_factors(self._h, _d(p), _d(res), _d(jac)), "lvio2d_eval_ground_factors")
        return res, jac

    def eval_edge_factor(self, tf12, weight, sqrt_info, pose_i, pose_j):
        """edge_factor (edge_factor.h:79-126): res[6], jac[6][12] over (p_i, q_i, p_j, q_j)."""
        res, jac = np.zeros(6), np.zeros((6, 12))
        a = [_f64(x).reshape(-1) for x in (tf12, sqrt_info, pose_i, pose_j)]
        self._check(self.lib.lvio2d_eval_edge_factor(self._h, _d(a[0]), float(weight), _d(a[1]), _d(a[2]), _d(a[3]), _d(res), _d(jac)),
                    "lvio2d_eval_edge_factor")
        return res, jac

    def pose_graph_solve(self, poses, edge_index, edge_tf, edge_weight, sqrt_info, ground_p=True, ground_q=True, fixed_pose=None):
        """keyframe_manager::solve (keyframe_manager.cpp:722-838) on the device: returns (poses [K][6], summary).
        fixed_pose: the constant key frame (-1: none); None = index1 of the first edge, the reference's seq_edges[0]."""
        x = np.array(poses, dtype=np.float64).reshape(-1, 6).copy()
        ei = np.ascontiguousarray(edge_index, dtype=np.int32).reshape(-1, 2)
        et = _f64(edge_tf).reshape(-1, 12)
        ew = _f64(edge_weight).reshape(-1)
        if len(et) != len(ei) or len(ew) != len(ei):
            raise ValueError("edge_index, edge_tf and edge_weight must describe the same number of edges")
        Jn = _f64(sqrt_info).reshape(-1)
        if Jn.size != 36:
            raise ValueError("sqrt_info is edge_noise::J, 6x6")
        summ = np.zeros(1, dtype=abi.SUMMARY_DTYPE)
        self._check(self.lib.lvio2d_pose_graph_solve(self._h, len(x), _d(x), len(ei), ei.ctypes.data_as(abi.c_int32_p), _d(et), _d(ew), _d(Jn),
                                                     int(bool(ground_p)), int(bool(ground_q)),
                                                     (int(ei[0, 0]) if len(ei) else -1) if fixed_pose is None else int(fixed_pose), summ.ctypes.data_as(C.c_void_p)),
                    "lvio2d_pose_graph_solve")
        return x, summ


# ------------------------------------------------------------------------------------------------------
# The reference's solver surface
class Submap:
    """lvio2d_submap: the reference sub-map and the one being spawned of `n_managers` independent laser managers, kept in
    device memory (reference state: laser_manager.h ref_submap_ptr / spawnning_ref_submap_ptr / last_add_tf /
    current_count; transitions: laser_manager.cpp:424-496).  Host-buffer flavour of the C entry points; the device-pointer
    flavour (on_device = 1) is reached through add_scan_device / match_device."""

    def __init__(self, ctx, line_params, n_managers, line_cap, filter_p, filter_q, n_accumulation):
        self._ctx, self.n_managers, self.line_cap = ctx, int(n_managers), int(line_cap)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.lvio2d_submap_create(ctx._h, self.n_managers, self.line_cap, C.byref(line_params), float(filter_p), float(filter_q),
                                                int(n_accumulation), C.byref(self._h)), "lvio2d_submap_create")

    def close(self):
        if self._h:
            self._ctx.lib.lvio2d_submap_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._ctx._check(self._ctx.lib.lvio2d_submap_reset(self._h), "lvio2d_submap_reset")

    def add_scan(self, n_lines, lines, pose):
        """n_lines [M] (negative: no scan for that manager), lines [M][max_lines][4], pose [M][6]."""
        n = np.ascontiguousarray(n_lines, dtype=np.int32).reshape(self.n_managers)
        l = np.ascontiguousarray(lines, dtype=np.float64).reshape(self.n_managers, -1, 4)
        s = np.ascontiguousarray(pose, dtype=np.float64).reshape(self.n_managers, 6)
        self._ctx._check(self._ctx.lib.lvio2d_submap_add_scan(self._h, l.shape[1], n.ctypes.data, l.ctypes.data, s.ctypes.data, 0), "lvio2d_submap_add_scan")

    def add_scan_device(self, max_lines, n_lines_ptr, lines_ptr, pose_ptr):
        self._ctx._check(self._ctx.lib.lvio2d_submap_add_scan(self._h, int(max_lines), n_lines_ptr, lines_ptr, pose_ptr, 1), "lvio2d_submap_add_scan")

    def get(self, which=0, want_lines=True):
        """Host copy: meta [M][4] (has reference, has spawning, current_count, last scan added), pose [M][6], n_lines [M], lines."""
        M = self.n_managers
        meta, pose, n = np.zeros((M, 4), np.int32), np.zeros((M, 6)), np.zeros(M, np.int32)
        lines = np.zeros((M, self.line_cap, 4)) if want_lines else None
        self._ctx._check(self._ctx.lib.lvio2d_submap_get(self._h, int(which), meta.ctypes.data_as(abi.c_int32_p), _d(pose), n.ctypes.data_as(abi.c_int32_p),
                                                         _d(lines) if lines is not None else abi.c_double_p()), "lvio2d_submap_get")
        return meta, pose, n, lines

    def device_pointers(self):
        """(meta, ref_pose, ref_n_lines, ref_lines) device addresses, slot 0 of the last three = the reference sub-map."""
        out = [C.c_void_p() for _ in range(4)]
        self._ctx._check(self._ctx.lib.lvio2d_submap_device(self._h, *[C.byref(o) for o in out]), "lvio2d_submap_device")
        return tuple(int(o.value or 0) for o in out)

    def match(self, n_lines2, lines2, pose2, kk=0, want_lines1=False):
        """laser_manager::match_with_ref for every manager: n_match [M], match [M][max_lines2][2] = (sub-map line, scan line);
        with want_lines1 also the matched sub-map lines' end points [M][max_lines2][4] and the sub-maps' poses [M][6]."""
        M = self.n_managers
        n = np.ascontiguousarray(n_lines2, dtype=np.int32).reshape(M)
        l = np.ascontiguousarray(lines2, dtype=np.float64).reshape(M, -1, 4)
        s = np.ascontiguousarray(pose2, dtype=np.float64).reshape(M, 6)
        nm, match = np.zeros(M, np.int32), np.zeros((M, l.shape[1], 2), np.int32)
        l1, rp = (np.zeros((M, l.shape[1], 4)), np.zeros((M, 6))) if want_lines1 else (None, None)
        self._ctx._check(self._ctx.lib.lvio2d_submap_match(self._h, int(kk), l.shape[1], n.ctypes.data, l.ctypes.data, s.ctypes.data, nm.ctypes.data,
                                                           match.ctypes.data, l1.ctypes.data if want_lines1 else None,
                                                           rp.ctypes.data if want_lines1 else None, 0), "lvio2d_submap_match")
        return (nm, match, l1, rp) if want_lines1 else (nm, match)

    def match_device(self, max_lines2, n_lines2_ptr, lines2_ptr, pose2_ptr, n_match_ptr, match_ptr, kk=0, lines1_ptr=None, ref_pose_ptr=None):
        self._ctx._check(self._ctx.lib.lvio2d_submap_match(self._h, int(kk), int(max_lines2), n_lines2_ptr, lines2_ptr, pose2_ptr, n_match_ptr,
                                                           match_ptr, lines1_ptr, ref_pose_ptr, 1), "lvio2d_submap_match")


class Line:
    """lvio_2d::line (reference src/trajectory/laser_type.h:13-21): end points in a laser frame, z = 0."""

    def __init__(self, p1, p2):
        self.p1 = np.asarray(p1, dtype=np.float64).reshape(3)
        self.p2 = np.asarray(p2, dtype=np.float64).reshape(3)


class LaserMatch:
    """lvio_2d::laser_match (laser_type.h:76-85): lines1[j] (reference submap, under pose p1/q1) matched with
    lines2[j] (this scan, under the frame's pose p2/q2)."""

    def __init__(self, lines1, lines2, p1, q1):
        assert len(lines1) == len(lines2)
        self.lines1, self.lines2 = list(lines1), list(lines2)
        self.p1, self.q1 = np.array(p1, dtype=np.float64), np.array(q1, dtype=np.float64)
        self.p2, self.q2 = np.zeros(3), np.zeros(3)


class FrameInfo:
    """lvio_2d::frame_info (reference src/trajectory/trajectory_type.h:9-75), laser frames only."""

    def __init__(self, time, p, q, v, bs, imu_observation_result=None, wheel_observation_result=None):
        self.time = float(time)
        self.p, self.q = np.array(p, dtype=np.float64), np.array(q, dtype=np.float64)
        self.v, self.bs = np.array(v, dtype=np.float64), np.array(bs, dtype=np.float64)
        self.imu_observation_result = imu_observation_result      # [466] blob of interval (i-1, i)
        self.wheel_observation_result = wheel_observation_result  # [15] blob
        self.laser_match = None
        self.is_key_frame = False
        self.sqrt_H = np.eye(6)

    def add_laser_match(self, match):
        self.laser_match = match


def _pair_weight(l1, l2):
    # laser_factor::sum (laser_factor.h:38-42)
    return float(np.sqrt(min(np.linalg.norm(l1.p1 - l1.p2), np.linalg.norm(l2.p1 - l2.p2)) / 2.0 / 0.02))


class Solver:
    """lvio_2d::solver (reference src/factor/solver.h:28-80) backed by the CUDA library.

    solve / init_solve / marginalization take the window as a list of FrameInfo and write the results back in
    place exactly where the reference does (solver.cpp:685-707 via Ceres, :804-814, :177-190, :401, :409-441).
    The camera path is not supported (`enable_camera: false` in every shipped config).
    """

    def __init__(self, params, fast_mode=False, ctx=None):
        self.params = params
        self.fast_mode = bool(fast_mode)
        # `ctx`: any object with Context's set_windows / solve / get_states / marginalize (tests pass a CPU-oracle
        # stand-in to produce the comparison trajectory; the product default is the CUDA library)
        self.ctx = ctx if ctx is not None else Context(params)
        self.has_linearized_block = False
        self.linearized_X = None
        self.linearized_jacobians = None
        self.linearized_residuals = None
        self.last_summary = None

    def close(self):
        self.ctx.close()

    # -- flatten a deque of frames into the C ABI batch
    def _batch(self, frames, topology, with_prior, laser_frames):
        n = len(frames)
        states = np.stack([np.concatenate([f.p, f.q, f.v, f.bs]) for f in frames])
        cmask = np.zeros(n, np.uint8)
        pts, pl, pw, lines, poff, loff = [], [], [], [], [0], [0]
        ref_frame = np.full(n, -1, np.int32)
        ref_pose = np.zeros((n, 6))
        for i, f in enumerate(frames):
            m = f.laser_match if i in laser_frames else None
            if m is not None and len(m.lines1) > 0:
                for j, (l1, l2) in enumerate(zip(m.lines1, m.lines2)):
                    w = _pair_weight(l1, l2)
                    pts += [l2.p1[:2], l2.p2[:2]]
                    pl += [j, j]
                    pw += [w, w]
                    lines.append(np.concatenate([l1.p1[:2], l1.p2[:2]]))
                if topology == "init":
                    ref_frame[i] = 0
                ref_pose[i, 0:3], ref_pose[i, 3:6] = m.p1, m.q1
            poff.append(len(pts))
            loff.append(len(lines))
        if topology == "tracking":
            # solver.cpp:787-794: every frame but the newest has p, q (and bs in fast_mode) held constant
            for i in range(n - 1):
                cmask[i] = abi.CONST_P | abi.CONST_Q | (abi.CONST_BS if self.fast_mode else 0)
        imu = np.stack([f.imu_observation_result for f in frames[1:]]) if n > 1 else None
        wheel = np.stack([f.wheel_observation_result for f in frames[1:]]) if n > 1 else None
        prior = with_prior and self.has_linearized_block and n >= 2
        return abi.HostBatch(
            1, n, n, (n - 2) if prior else -1,
            states=states, const_mask=cmask, point_offset=np.array(poff, np.int64),
            points=np.array(pts, dtype=np.float64).reshape(-1, 2), point_line=np.array(pl, np.int32),
            point_weight=np.array(pw, dtype=np.float64), line_offset=np.array(loff, np.int64),
            lines=np.array(lines, dtype=np.float64).reshape(-1, 4), ref_frame=ref_frame, ref_pose=ref_pose,
            imu=imu, wheel=wheel,
            prior_X0=self.linearized_X if prior else None, prior_J=self.linearized_jacobians if prior else None)

    def _write_back(self, frames, states):
        for f, s in zip(frames, states.reshape(len(frames), 15)):
            f.p[:], f.q[:], f.v[:], f.bs[:] = s[0:3], s[3:6], s[6:9], s[9:15]

    def solve(self, frame_infos, feature_infos=None):
        """solver::solve (solver.cpp:631-820): laser factors of the newest frame against the constant reference
        pose, IMU/wheel chain, ground factors x n, marginalisation prior on frame n-2 unless fast_mode."""
        n = len(frame_infos)
        hb = self._batch(frame_infos, "tracking", with_prior=not self.fast_mode, laser_frames={n - 1})
        self.ctx.set_windows(hb)
        if self.fast_mode:
            self.ctx.set_max_iterations(10)          # solver.cpp:800-801
        self.last_summary = self.ctx.solve()
        self._write_back(frame_infos, self.ctx.get_states())
        m = frame_infos[-1].laser_match
        if m is not None:
            m.p2, m.q2 = frame_infos[-1].p.copy(), frame_infos[-1].q.copy()  # solver.cpp:804-814

    def init_solve(self, frame_infos, feature_infos=None):
        """solver::init_solve (solver.cpp:171-195): laser factors between frame 0 and every laser frame, nothing
        constant, then the laser_match poses are refreshed."""
        n = len(frame_infos)
        hb = self._batch(frame_infos, "init", with_prior=False, laser_frames=set(range(1, n)))
        self.ctx.set_windows(hb)
        # do_init_solve leaves max_num_iterations at Ceres' default 50 even in fast_mode (solver.cpp:161-168)
        before = int(self.ctx.params.max_iters)
        if self.fast_mode:
            self.ctx.set_max_iterations(50)
        try:
            self.last_summary = self.ctx.solve()
        finally:
            if self.fast_mode:
                self.ctx.set_max_iterations(before)
        self._write_back(frame_infos, self.ctx.get_states())
        for f in frame_infos:
            if f.laser_match is not None:
                f.laser_match.p1, f.laser_match.q1 = frame_infos[0].p.copy(), frame_infos[0].q.copy()
                f.laser_match.p2, f.laser_match.q2 = f.p.copy(), f.q.copy()

    def marginalization(self, frame_infos, feature_infos=None):
        """solver::marginalization (solver.cpp:257-442): early return in fast_mode; otherwise linearise every
        factor at the current states, Schur-complement onto the newest frame, keep the sqrt-information prior."""
        if self.fast_mode:
            return
        n = len(frame_infos)
        hb = self._batch(frame_infos, "marg", with_prior=True, laser_frames=set(range(n)))
        self.ctx.set_windows(hb)
        X0, J, r = self.ctx.marginalize()
        self.linearized_X, self.linearized_jacobians, self.linearized_residuals = X0[0], J[0], r[0]
        frame_infos[-1].sqrt_H = J[0][0:6, 0:6].copy()  # solver.cpp:401
        self.has_linearized_block = True
