"""ctypes mirror of include/lvio2d.h (the C ABI of the B200 front-end solver).

Pure declarations: the structs are shared by the product host code (solver.py) and by the test-side
checker in tests/ so that both sides receive byte-identical inputs.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 1
STATE_DIM = 15
IMU_BLOB = 466
WHEEL_BLOB = 15
LASER_BLOCK = 28

OK = 0
ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_NO_WINDOW, ERR_DOMAIN, ERR_ALLOC = -1, -2, -3, -4, -5, -6

CONST_P, CONST_Q, CONST_V, CONST_BS = 1, 2, 4, 8
ASSOC_FIXED, ASSOC_NEAREST = 0, 1

TERM_NO_CONVERGENCE = 0
TERM_CONVERGENCE_FUNCTION = 1
TERM_CONVERGENCE_PARAMETER = 2
TERM_CONVERGENCE_GRADIENT = 3
TERM_CONVERGENCE_RADIUS = 4
TERM_FAILURE = 5

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class Params(C.Structure):
    """lvio2d_params (include/lvio2d.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("T_imu_to_laser", C.c_double * 12),
        ("T_imu_to_wheel", C.c_double * 12),
        ("g", C.c_double),
        ("line_to_line_sigma", C.c_double),
        ("manifold_p_sigma", C.c_double),
        ("manifold_q_sigma", C.c_double),
        ("imu_noise_acc_sigma", C.c_double * 3),
        ("imu_bias_acc_sigma", C.c_double * 3),
        ("imu_noise_gyro_sigma", C.c_double * 3),
        ("imu_bias_gyro_sigma", C.c_double * 3),
        ("wheel_sigma", C.c_double * 3),
        ("max_iters", C.c_int32),
        ("assoc_mode", C.c_int32),
        ("huber_delta", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double),
        ("assoc_gate", C.c_double),
        ("assoc_max_dist", C.c_double),
    ]


class LineParams(C.Structure):
    """lvio2d_line_params (include/lvio2d.h)."""

    _fields_ = [
        ("line_continuous_threshold", C.c_double),
        ("line_max_tolerance_angle_deg", C.c_double),
        ("line_max_dis", C.c_double),
        ("line_min_len", C.c_double),
        ("laser_resolution", C.c_double),
        ("w_laser_each_scan", C.c_double),
        ("h_laser_each_scan", C.c_double),
    ]


class ScanHeader(C.Structure):
    """lvio2d_scan_header (include/lvio2d.h)."""

    _fields_ = [
        ("angle_min", C.c_float),
        ("angle_increment", C.c_float),
        ("time_increment", C.c_float),
        ("reserved", C.c_float),
        ("stamp", C.c_double),
        ("linear", C.c_double * 3),
        ("angular", C.c_double * 3),
    ]


SCAN_HEADER_DTYPE = np.dtype([("angle_min", np.float32), ("angle_increment", np.float32), ("time_increment", np.float32),
                              ("reserved", np.float32), ("stamp", np.float64), ("linear", np.float64, 3), ("angular", np.float64, 3)])
assert SCAN_HEADER_DTYPE.itemsize == C.sizeof(ScanHeader)


class WindowBatch(C.Structure):
    """lvio2d_window_batch (include/lvio2d.h)."""

    _fields_ = [
        ("n_windows", C.c_int32),
        ("n_frames", C.c_int32),
        ("states", c_double_p),
        ("const_mask", c_uint8_p),
        ("point_offset", c_int64_p),
        ("points", c_double_p),
        ("point_line", c_int32_p),
        ("point_weight", c_double_p),
        ("line_offset", c_int64_p),
        ("lines", c_double_p),
        ("ref_frame", c_int32_p),
        ("ref_pose", c_double_p),
        ("imu", c_double_p),
        ("wheel", c_double_p),
        ("ground_multiplicity", C.c_int32),
        ("prior_frame", C.c_int32),
        ("prior_X0", c_double_p),
        ("prior_J", c_double_p),
    ]


IMU_COMPACT = 190


class ScanWireStruct(C.Structure):
    """lvio2d_scan_wire (include/lvio2d.h)."""

    _fields_ = [("n_beams", C.c_int32), ("ranges", C.POINTER(C.c_float)), ("angle", C.POINTER(C.c_float)), ("beam_line", C.POINTER(C.c_uint16)),
                ("imu_compact", c_double_p), ("shared_lines", C.c_int32), ("reserved", C.c_int32), ("beam_line8", C.POINTER(C.c_uint8))]


class ScanWire:
    """Compact wire encoding of a beam-mode batch's laser input: float32 ranges [F][n_beams], float32 (angle_min,
    angle_increment) [F][2], uint16 line index per beam [F][n_beams] (0xFFFF = none).  `points()` is the host statement
    of what the device rebuilds (convert::laser_to_point_times without the 1 cm thinning, src/utilies/common.cpp:6-24)."""

    NONE = 0xFFFF

    def __init__(self, ranges, angle, beam_line, imu_compact=None, shared_lines=False):
        self.shared_lines = bool(shared_lines)   # the batch's lines / line_offset are per WINDOW (one sub-map per window)
        self.beam_line8 = None                    # optional 8-bit flavour of beam_line (narrow())
        self.imu_compact = None if imu_compact is None else np.ascontiguousarray(imu_compact, dtype=np.float64).reshape(-1, IMU_COMPACT)
        self.ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        self.angle = np.ascontiguousarray(angle, dtype=np.float32).reshape(-1, 2)
        self.beam_line = np.ascontiguousarray(beam_line, dtype=np.uint16)
        assert self.ranges.ndim == 2 and self.beam_line.shape == self.ranges.shape and len(self.angle) == len(self.ranges)
        self.n_beams = int(self.ranges.shape[1])

    def struct(self):
        s = ScanWireStruct()
        s.n_beams = self.n_beams
        s.ranges = self.ranges.ctypes.data_as(C.POINTER(C.c_float))
        s.angle = self.angle.ctypes.data_as(C.POINTER(C.c_float))
        s.beam_line = self.beam_line.ctypes.data_as(C.POINTER(C.c_uint16))
        s.imu_compact = ptr(self.imu_compact, c_double_p)
        s.shared_lines, s.reserved = int(self.shared_lines), 0
        s.beam_line8 = self.beam_line8.ctypes.data_as(C.POINTER(C.c_uint8)) if self.beam_line8 is not None else C.POINTER(C.c_uint8)()
        return s

    def narrow(self):
        """8-bit line indices (0xFF = none) when every index fits: 5 instead of 6 bytes per beam.  Returns True when used."""
        valid = self.beam_line != self.NONE
        if valid.any() and int(self.beam_line[valid].max()) >= 255:
            return False
        self.beam_line8 = np.where(valid, self.beam_line, 255).astype(np.uint8)
        return True

    def nbytes(self):
        idx = self.beam_line.nbytes if self.beam_line8 is None else self.beam_line8.nbytes
        return self.ranges.nbytes + self.angle.nbytes + idx + (0 if self.imu_compact is None else self.imu_compact.nbytes)

    @staticmethod
    def compact_imu(blobs):
        """[k][466] imu_preint_result blobs -> [k][190]: the entries imu_factor::operator() reads (imu_factor.h:52-86)."""
        b = np.asarray(blobs, dtype=np.float64).reshape(-1, IMU_BLOB)
        J = b[:, 15:240].reshape(-1, 15, 15)
        S = b[:, 240:465].reshape(-1, 15, 15)
        iu = np.triu_indices(15)
        return np.ascontiguousarray(np.concatenate([b[:, 0:15], J[:, 0:9, 9:15].reshape(len(b), 54), S[:, iu[0], iu[1]], b[:, 465:466]], axis=1))

    def points(self):
        """-> points [F * n_beams][2] float64, point_line [F * n_beams] int32 (-1 = no part), point_offset [F + 1]"""
        k = np.arange(self.n_beams, dtype=np.float32)[None, :]
        ang = (self.angle[:, 0:1] + k * self.angle[:, 1:2]).astype(np.float32)      # float32 products and sums, rounded each
        a64, r = ang.astype(np.float64), self.ranges.astype(np.float64)
        valid = np.isfinite(self.ranges) & (r > 0.1)
        pts = np.stack([np.cos(a64) * r, np.sin(a64) * r], axis=-1)
        pts[~valid] = 0.0
        line = np.where(valid & (self.beam_line != self.NONE), self.beam_line.astype(np.int32), -1).astype(np.int32)
        off = np.arange(len(self.ranges) + 1, dtype=np.int64) * self.n_beams
        return pts.reshape(-1, 2), line.reshape(-1), off

    @classmethod
    def from_points(cls, hb, n_beams, angle_min, angle_increment):
        """Wire form of a HostBatch whose frames hold exactly one point per beam, in beam order: the range is |p|, the
        direction snaps to the float32 beam grid (synthetic data: nothing is lost but sub-float32 digits)."""
        pts = hb["points"].reshape(-1, n_beams, 2)
        F = len(pts)
        assert F == hb.n_windows * hb.n_frames and np.array_equal(np.diff(hb["point_offset"]), np.full(F, n_beams))
        rng = np.hypot(pts[..., 0], pts[..., 1]).astype(np.float32)
        pl = hb["point_line"].reshape(F, n_beams)
        assert pl.max() < cls.NONE
        bl = np.where(pl < 0, cls.NONE, pl).astype(np.uint16)
        ang = np.tile(np.array([[angle_min, angle_increment]], dtype=np.float32), (F, 1))
        return cls(rng, ang, bl)


class Summary(C.Structure):
    """lvio2d_summary (include/lvio2d.h)."""

    _fields_ = [
        ("iterations", C.c_int32),
        ("termination", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("final_radius", C.c_double),
        ("reserved", C.c_double),
    ]


SUMMARY_DTYPE = np.dtype(
    [
        ("iterations", np.int32),
        ("termination", np.int32),
        ("num_successful_steps", np.int32),
        ("num_unsuccessful_steps", np.int32),
        ("initial_cost", np.float64),
        ("final_cost", np.float64),
        ("final_radius", np.float64),
        ("reserved", np.float64),
    ]
)
assert SUMMARY_DTYPE.itemsize == C.sizeof(Summary)

_PTR_FIELDS = {
    "states": (np.float64, c_double_p),
    "const_mask": (np.uint8, c_uint8_p),
    "point_offset": (np.int64, c_int64_p),
    "points": (np.float64, c_double_p),
    "point_line": (np.int32, c_int32_p),
    "point_weight": (np.float64, c_double_p),
    "line_offset": (np.int64, c_int64_p),
    "lines": (np.float64, c_double_p),
    "ref_frame": (np.int32, c_int32_p),
    "ref_pose": (np.float64, c_double_p),
    "imu": (np.float64, c_double_p),
    "wheel": (np.float64, c_double_p),
    "prior_X0": (np.float64, c_double_p),
    "prior_J": (np.float64, c_double_p),
}


def ptr(arr, ctype):
    return arr.ctypes.data_as(ctype) if arr is not None else ctype()


class HostBatch:
    """A window batch held as contiguous numpy arrays + the ctypes struct that points into them.

    `fields` maps lvio2d_window_batch member names to arrays (or None for NULL pointers).
    """

    def __init__(self, n_windows, n_frames, ground_multiplicity, prior_frame=-1, **fields):
        self.n_windows = int(n_windows)
        self.n_frames = int(n_frames)
        self.ground_multiplicity = int(ground_multiplicity)
        self.prior_frame = int(prior_frame)
        self.arrays = {}
        for name, (dtype, _) in _PTR_FIELDS.items():
            a = fields.get(name)
            self.arrays[name] = None if a is None else np.ascontiguousarray(a, dtype=dtype)
        unknown = set(fields) - set(_PTR_FIELDS)
        if unknown:
            raise TypeError(f"unknown batch fields {sorted(unknown)}")
        self._check()

    def _check(self):
        B, n = self.n_windows, self.n_frames
        a = self.arrays
        assert a["states"] is not None and a["states"].size == B * n * STATE_DIM
        if a["const_mask"] is not None:
            assert a["const_mask"].size == B * n
        if a["point_offset"] is not None:
            assert a["point_offset"].size == B * n + 1
            N = int(a["point_offset"][-1])
            assert a["points"].size == 2 * N and a["point_line"].size == N
            assert a["line_offset"].size == B * n + 1
            assert a["lines"].size == 4 * int(a["line_offset"][-1])
            if a["point_weight"] is not None:
                assert a["point_weight"].size == N
            if a["ref_frame"] is not None:
                assert a["ref_frame"].size == B * n
            assert a["ref_pose"] is not None and a["ref_pose"].size == B * n * 6
        if a["imu"] is not None:
            assert a["imu"].size == B * (n - 1) * IMU_BLOB
        if a["wheel"] is not None:
            assert a["wheel"].size == B * (n - 1) * WHEEL_BLOB
        if self.prior_frame >= 0:
            assert a["prior_X0"].size == B * 15 and a["prior_J"].size == B * 225

    def __getitem__(self, name):
        return self.arrays[name]

    @property
    def n_points(self):
        po = self.arrays["point_offset"]
        return 0 if po is None else int(po[-1])

    def struct(self):
        s = WindowBatch()
        s.n_windows, s.n_frames = self.n_windows, self.n_frames
        s.ground_multiplicity, s.prior_frame = self.ground_multiplicity, self.prior_frame
        for name, (_, ctype) in _PTR_FIELDS.items():
            setattr(s, name, ptr(self.arrays[name], ctype))
        return s

    def replace(self, **fields):
        merged = dict(self.arrays)
        merged.update(fields)
        kw = dict(n_windows=self.n_windows, n_frames=self.n_frames, ground_multiplicity=self.ground_multiplicity,
                  prior_frame=self.prior_frame)
        for k in ("n_windows", "n_frames", "ground_multiplicity", "prior_frame"):
            if k in merged:
                kw[k] = merged.pop(k)
        return HostBatch(**kw, **merged)

    def nbytes(self):
        return sum(a.nbytes for a in self.arrays.values() if a is not None)
