"""Deterministic synthetic sliding windows for the BASELINE.json configs.

A window is generated from raw "sensor" streams — a planar robot driving through a 10 m x 8 m room
with 96 interior wall segments, a 2-D lidar ray-cast against the walls, 200 Hz IMU samples and wheel
odometry steps — using the corridor extrinsics/noise values.  The streams are turned into the
solver's inputs the way the reference front-end does it: scans become points (+ a line
correspondence per point) against a <=100-segment local map expressed in a reference laser frame
(laser_match, reference src/trajectory/laser_type.h:76-85), IMU/wheel streams go through the
preintegrators (src/factor/imu_preintegraption.h, wheel_odom_preintegration.h).

Everything here is host-side numpy; nothing is imported from oracle/.  RNG: numpy MT19937, seed
42 + window index.
"""
import math

import numpy as np

from . import abi
from .params import corridor_params, params_T


# ----------------------------------------------------------------------------- small SO(3) helpers
def hat(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def exp_so3(v):
    """Rodrigues formula (same rotation as lie::exp_so3, reference src/utilies/common.h:137-146)."""
    v = np.asarray(v, dtype=np.float64)
    th = float(np.linalg.norm(v))
    K = hat(v)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + (math.sin(th) / th) * K + ((1.0 - math.cos(th)) / (th * th)) * (K @ K)


def log_so3(R):
    """Angle-axis vector with |v| <= pi (same rotation as lie::log_SO3, common.h:148-163)."""
    tr = max(-1.0, min(3.0, float(np.trace(R))))
    c = 0.5 * (tr - 1.0)
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = float(np.linalg.norm(w))
    th = math.atan2(s, c)
    if s > 1e-9:
        return w * (th / s)
    if c > 0:
        return w
    # theta ~ pi: take the axis from the symmetric part
    A = 0.5 * (R + np.eye(3))
    k = int(np.argmax(np.diag(A)))
    ax = A[:, k] / math.sqrt(max(A[k, k], 1e-300))
    return ax * th


def rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


# ----------------------------------------------------------------------------- world + trajectory
def make_world(rng, n_segments=100, width=10.0, height=8.0):
    """4 walls + (n_segments - 4) interior segments of length U(0.3, 2.0). Returns [L][4] = x1 y1 x2 y2."""
    segs = [
        [0.0, 0.0, width, 0.0],
        [width, 0.0, width, height],
        [width, height, 0.0, height],
        [0.0, height, 0.0, 0.0],
    ]
    for _ in range(max(0, n_segments - 4)):
        length = rng.uniform(0.3, 2.0)
        ang = rng.uniform(0.0, math.pi)
        cx = rng.uniform(0.5, width - 0.5)
        cy = rng.uniform(0.5, height - 0.5)
        dx, dy = 0.5 * length * math.cos(ang), 0.5 * length * math.sin(ang)
        segs.append([cx - dx, cy - dy, cx + dx, cy + dy])
    return np.array(segs[:n_segments], dtype=np.float64)


class Trajectory:
    """Smooth planar base (wheel-frame) trajectory: speed 0.5..1.0 m/s, |yaw rate| <= 0.5 rad/s, so that
    frames 0.1 s apart move 0.05..0.10 m and turn <= 0.05 rad (cf. key_frame_p_motion_threshold,
    config/corridor.yaml:94)."""

    def __init__(self, rng, duration, width=10.0, height=8.0):
        self.ph = rng.uniform(0.0, 2.0 * math.pi, size=4)
        self.psi0 = rng.uniform(-math.pi, math.pi)
        self.x0 = np.array([width / 2 + rng.uniform(-1.0, 1.0), height / 2 + rng.uniform(-1.0, 1.0)])
        self.h = 1e-4
        t = np.arange(0.0, duration + 0.02, self.h)
        v, psi = self.speed(t), self.yaw(t)
        vx, vy = v * np.cos(psi), v * np.sin(psi)
        self.x = self.x0[0] + np.concatenate([[0.0], np.cumsum(0.5 * (vx[1:] + vx[:-1]) * self.h)])
        self.y = self.x0[1] + np.concatenate([[0.0], np.cumsum(0.5 * (vy[1:] + vy[:-1]) * self.h)])
        self.t = t

    def speed(self, t):
        return 0.75 + 0.2 * np.sin(0.9 * t + self.ph[0]) + 0.05 * np.sin(2.3 * t + self.ph[1])

    def accel(self, t):
        return 0.18 * np.cos(0.9 * t + self.ph[0]) + 0.115 * np.cos(2.3 * t + self.ph[1])

    def yaw_rate(self, t):
        return 0.35 * np.sin(0.7 * t + self.ph[2]) + 0.15 * np.sin(1.9 * t + self.ph[3])

    def yaw_acc(self, t):
        return 0.245 * np.cos(0.7 * t + self.ph[2]) + 0.285 * np.cos(1.9 * t + self.ph[3])

    def yaw(self, t):
        return (self.psi0 - 0.5 * (np.cos(0.7 * t + self.ph[2]) - math.cos(self.ph[2]))
                - (0.15 / 1.9) * (np.cos(1.9 * t + self.ph[3]) - math.cos(self.ph[3])))

    def pos(self, t):
        return np.array([np.interp(t, self.t, self.x), np.interp(t, self.t, self.y)])


class SensorWindow:
    """Raw streams + solver inputs of one window (everything but the preintegrated blobs)."""


def _imu_pose_from_base(R_wb, p_wb, T_io):
    """T_w_imu = T_w_base * T_imu_to_wheel^-1 (tf_w_o = tf_w_i * T_i_w, reference ground_factor.h:42)."""
    R_io, t_io = T_io[:, :3], T_io[:, 3]
    R_wi = R_wb @ R_io.T
    p_wi = p_wb - R_wi @ t_io
    return R_wi, p_wi


def make_window(seed=42, n_frames=30, beams=1081, fov_deg=270.0, n_segments=100, topology="tracking",
                mode="beam", params=None, frame_dt=0.1, imu_rate=200.0, wheel_substeps=2, range_sigma=0.01,
                max_range=30.0, with_prior=True, guess_sigma_p=0.02, guess_sigma_q=0.01):
    """One synthetic window.

    topology "tracking": the tracking topology of solver::solve (reference src/factor/solver.cpp:631-794)
        generalised to n frames: frame 0 (p,q) constant, every frame's points matched against a local map
        under an external constant reference pose, IMU/wheel chain, ground factors x n, prior on frame 0.
    topology "init": do_init_solve (solver.cpp:50-169): laser terms between frame 0 and frame i with both
        poses free, nothing constant, no prior.
    mode "beam": every valid beam is a point with weight 1 (x 1/line_to_line_sigma inside the factor).
    mode "segment": the reference's laser_factor input: 2 end points per matched line pair with the pair
        weight sqrt(min(len1,len2)/2/0.02) (laser_factor.h:38-42).
    """
    P = params if params is not None else corridor_params()
    T_il, T_io = params_T(P, "T_imu_to_laser"), params_T(P, "T_imu_to_wheel")
    rng = np.random.Generator(np.random.MT19937(int(seed)))
    world = make_world(rng, n_segments)
    n = int(n_frames)
    traj = Trajectory(rng, duration=n * frame_dt)
    tf = np.arange(n) * frame_dt

    # ---- true states at the frame times.  Planar motion + a small per-frame roll/pitch/z perturbation so
    # that no ground residual is exactly zero (SURVEY.md §7 hard parts)
    r_bi = -T_io[:, :3].T @ T_io[:, 3]  # IMU origin in the base frame
    truth = np.zeros((n, 15))
    R_wl, t_wl = [], []
    ba_true = rng.normal(0.0, 0.02, size=3)
    bw_true = rng.normal(0.0, 0.002, size=3)
    for k in range(n):
        psi = float(traj.yaw(tf[k]))
        roll, pitch, z = rng.normal(0.0, 1e-3, size=3)
        R_wb = rot_z(psi) @ rot_y(pitch) @ rot_x(roll)
        p_wb = np.array([*traj.pos(tf[k]), z])
        R_wi, p_wi = _imu_pose_from_base(R_wb, p_wb, T_io)
        v_b = float(traj.speed(tf[k])) * np.array([math.cos(psi), math.sin(psi), 0.0])
        v_i = v_b + np.cross(np.array([0.0, 0.0, float(traj.yaw_rate(tf[k]))]), rot_z(psi) @ r_bi)
        truth[k, 0:3], truth[k, 3:6], truth[k, 6:9] = p_wi, log_so3(R_wi), v_i
        truth[k, 9:12], truth[k, 12:15] = ba_true, bw_true
        R_wl.append(R_wi @ T_il[:, :3])
        t_wl.append(R_wi @ T_il[:, 3] + p_wi)

    # ---- scans: ray-cast every beam against the world in the (flattened) world xy-plane
    if fov_deg < 360.0:
        ang = np.deg2rad(np.linspace(-fov_deg / 2.0, fov_deg / 2.0, beams))
    else:
        ang = np.deg2rad(np.arange(beams) * (360.0 / beams) - 180.0)
    ca, sa = np.cos(ang), np.sin(ang)
    s1, e = world[:, 0:2], world[:, 2:4] - world[:, 0:2]
    pts_per_frame, line_per_frame, w_per_frame = [], [], []
    for k in range(n):
        if topology == "init" and k == 0:
            pts_per_frame.append(np.zeros((0, 2)))
            line_per_frame.append(np.zeros(0, np.int32))
            w_per_frame.append(np.zeros(0) if mode == "segment" else None)
            continue
        A = R_wl[k][:2, :2]
        o = t_wl[k][:2]
        d = np.stack([A[0, 0] * ca + A[0, 1] * sa, A[1, 0] * ca + A[1, 1] * sa], axis=1)  # [beams][2]
        so = s1 - o  # [L][2]
        den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            rho = (so[None, :, 0] * e[None, :, 1] - so[None, :, 1] * e[None, :, 0]) / den
            u = (so[None, :, 0] * d[:, None, 1] - so[None, :, 1] * d[:, None, 0]) / den
        ok = (np.abs(den) > 1e-12) & (rho > 0.05) & (u >= 0.0) & (u <= 1.0)
        rho = np.where(ok, rho, np.inf)
        hit = np.argmin(rho, axis=1)
        r = rho[np.arange(beams), hit]
        valid = np.isfinite(r) & (r < max_range)
        r_noisy = r + rng.normal(0.0, range_sigma, size=beams)
        pts = np.stack([r_noisy * ca, r_noisy * sa], axis=1)[valid]
        idx = hit[valid].astype(np.int32)
        if mode == "segment":
            # consecutive beams on the same wall -> one current-scan segment; pair = (wall, segment)
            seg_pts, seg_idx, seg_w = [], [], []
            start = 0
            for b in range(1, len(idx) + 1):
                if b == len(idx) or idx[b] != idx[start]:
                    if b - start >= 5:
                        c1, c2 = pts[start], pts[b - 1]
                        len2 = float(np.linalg.norm(c1 - c2))
                        wl = world[idx[start]]
                        len1 = float(np.linalg.norm(wl[0:2] - wl[2:4]))
                        wgt = math.sqrt(min(len1, len2) / 2.0 / 0.02)
                        seg_pts += [c1, c2]
                        seg_idx += [idx[start]] * 2
                        seg_w += [wgt] * 2
                    start = b
            pts = np.array(seg_pts).reshape(-1, 2)
            idx = np.array(seg_idx, dtype=np.int32)
            w_per_frame.append(np.array(seg_w))
        else:
            w_per_frame.append(None)
        pts_per_frame.append(pts)
        line_per_frame.append(idx)

    # ---- local maps: world segments expressed in the reference laser frame (z = 0 there)
    ref_frame = np.full(n, -1, np.int32)
    ref_pose = np.zeros((n, 6))
    lines_per_frame = []
    for k in range(n):
        if topology == "init":
            ref_frame[k] = 0 if k > 0 else -1
        ref_pose[k] = truth[0, 0:6]
        Rr, tr = R_wl[0], t_wl[0]
        Ainv = np.linalg.inv(Rr[:2, :2])
        a1 = (world[:, 0:2] - tr[:2]) @ Ainv.T
        a2 = (world[:, 2:4] - tr[:2]) @ Ainv.T
        lines_per_frame.append(np.concatenate([a1, a2], axis=1))

    # ---- initial guess: truth + N(0, .02 m) / N(0, .01 rad); frame 0 pose exact in the tracking topology
    states0 = truth.copy()
    states0[:, 0:3] += rng.normal(0.0, guess_sigma_p, size=(n, 3))
    states0[:, 3:6] += rng.normal(0.0, guess_sigma_q, size=(n, 3))
    states0[:, 6:9] += rng.normal(0.0, 0.05, size=(n, 3))
    states0[:, 9:12] += rng.normal(0.0, 0.005, size=(n, 3))
    states0[:, 12:15] += rng.normal(0.0, 0.0005, size=(n, 3))
    const_mask = np.zeros(n, np.uint8)
    if topology == "tracking":
        states0[0, 0:6] = truth[0, 0:6]
        const_mask[0] = abi.CONST_P | abi.CONST_Q

    # ---- IMU samples: row = (dt, acc, gyro) integrated by one update(dt) (imu_preintegraption.h:170-208)
    dt_imu = 1.0 / imu_rate
    per = int(round(frame_dt * imu_rate))
    sig_a = np.array(list(P.imu_noise_acc_sigma))
    sig_w = np.array(list(P.imu_noise_gyro_sigma))
    R_io = T_io[:, :3]
    zhat = np.array([0.0, 0.0, 1.0])
    imu_rows, imu_off = [], [0]
    for k in range(1, n):
        for m in range(per):
            tm = tf[k - 1] + (m + 0.5) * dt_imu  # mid-step value
            psi, w_z, al = float(traj.yaw(tm)), float(traj.yaw_rate(tm)), float(traj.yaw_acc(tm))
            v, vd = float(traj.speed(tm)), float(traj.accel(tm))
            a_b = np.array([vd * math.cos(psi) - v * w_z * math.sin(psi), vd * math.sin(psi) + v * w_z * math.cos(psi), 0.0])
            Rr = rot_z(psi) @ r_bi
            a_i = a_b + al * np.cross(zhat, Rr) + w_z * w_z * np.cross(zhat, np.cross(zhat, Rr))
            R_wi = rot_z(psi) @ R_io.T
            f = R_wi.T @ (a_i + np.array([0.0, 0.0, P.g]))
            gyro = R_io @ np.array([0.0, 0.0, w_z])
            acc_m = f + ba_true + rng.normal(0.0, 1.0, 3) * sig_a
            gyr_m = gyro + bw_true + rng.normal(0.0, 1.0, 3) * sig_w
            imu_rows.append([dt_imu, *acc_m, *gyr_m])
        imu_off.append(len(imu_rows))
    bias0 = states0[:-1, 9:15].copy() if n > 1 else np.zeros((0, 6))

    # ---- wheel odometry steps: row = (dt, v, omega) for one update_by_v (wheel_odom_preintegration.h:141-152)
    wheel_rows, wheel_off = [], [0]
    sub_dt = frame_dt / wheel_substeps
    for k in range(1, n):
        for m in range(wheel_substeps):
            t0, t1 = tf[k - 1] + m * sub_dt, tf[k - 1] + (m + 1) * sub_dt
            R0, R1 = rot_z(float(traj.yaw(t0))), rot_z(float(traj.yaw(t1)))
            dp = R0.T @ np.array([*(traj.pos(t1) - traj.pos(t0)), 0.0])
            dq = log_so3(R0.T @ R1)
            vv = dp / sub_dt * (1.0 + rng.normal(0.0, 0.01)) + rng.normal(0.0, 1e-4, 3)
            ww = dq / sub_dt * (1.0 + rng.normal(0.0, 0.01)) + rng.normal(0.0, 1e-4, 3)
            wheel_rows.append([sub_dt, *vv, *ww])
        wheel_off.append(len(wheel_rows))

    W = SensorWindow()
    W.seed, W.n_frames, W.topology, W.mode = seed, n, topology, mode
    W.world, W.truth, W.states0, W.const_mask = world, truth, states0, const_mask
    W.points, W.point_line, W.point_weight, W.lines = pts_per_frame, line_per_frame, w_per_frame, lines_per_frame
    W.ref_frame, W.ref_pose = ref_frame, ref_pose
    W.imu_samples = np.array(imu_rows, dtype=np.float64).reshape(-1, 7)
    W.imu_offset = np.array(imu_off, dtype=np.int64)
    W.bias0 = bias0
    W.wheel_steps = np.array(wheel_rows, dtype=np.float64).reshape(-1, 7)
    W.wheel_offset = np.array(wheel_off, dtype=np.int64)
    W.ground_multiplicity = n
    if with_prior and topology == "tracking":
        # sqrt-information prior on frame 0 (stands in for the marginalisation prior, solver.cpp:744-785)
        W.prior_frame = 0
        W.prior_X0 = truth[0].copy()
        sig = np.array([0.01] * 3 + [0.01] * 3 + [0.1] * 3 + [0.05] * 3 + [0.005] * 3)
        W.prior_J = np.diag(1.0 / sig)
    else:
        W.prior_frame, W.prior_X0, W.prior_J = -1, None, None
    return W


class SensorBatch:
    """B windows concatenated into the flat layout of lvio2d_window_batch, plus the raw IMU / wheel
    streams that still have to go through a preintegrator."""

    def __init__(self, windows):
        self.windows = windows
        B, n = len(windows), windows[0].n_frames
        self.n_windows, self.n_frames = B, n
        self.states = np.concatenate([w.states0 for w in windows]).reshape(B * n, 15)
        self.truth = np.concatenate([w.truth for w in windows]).reshape(B * n, 15)
        self.const_mask = np.concatenate([w.const_mask for w in windows])
        pts = [p for w in windows for p in w.points]
        self.point_offset = np.concatenate([[0], np.cumsum([len(p) for p in pts])]).astype(np.int64)
        self.points = np.concatenate(pts).reshape(-1, 2) if pts else np.zeros((0, 2))
        self.point_line = np.concatenate([l for w in windows for l in w.point_line]).astype(np.int32)
        if windows[0].mode == "segment":
            self.point_weight = np.concatenate([x for w in windows for x in w.point_weight if x is not None])
        else:
            self.point_weight = None
        lines = [l for w in windows for l in w.lines]
        self.line_offset = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.int64)
        self.lines = np.concatenate(lines).reshape(-1, 4)
        self.ref_frame = np.concatenate([w.ref_frame for w in windows]).astype(np.int32)
        self.ref_pose = np.concatenate([w.ref_pose for w in windows]).reshape(B * n, 6)
        self.imu_samples = np.concatenate([w.imu_samples for w in windows]).reshape(-1, 7)
        off, acc = [0], 0
        for w in windows:
            off += (acc + w.imu_offset[1:]).tolist()
            acc += int(w.imu_offset[-1])
        self.imu_offset = np.array(off, dtype=np.int64)
        self.bias0 = np.concatenate([w.bias0 for w in windows]).reshape(-1, 6)
        self.wheel_steps = np.concatenate([w.wheel_steps for w in windows]).reshape(-1, 7)
        off, acc = [0], 0
        for w in windows:
            off += (acc + w.wheel_offset[1:]).tolist()
            acc += int(w.wheel_offset[-1])
        self.wheel_offset = np.array(off, dtype=np.int64)
        self.ground_multiplicity = windows[0].ground_multiplicity
        self.prior_frame = windows[0].prior_frame
        if self.prior_frame >= 0:
            self.prior_X0 = np.stack([w.prior_X0 for w in windows])
            self.prior_J = np.stack([w.prior_J for w in windows])
        else:
            self.prior_X0 = self.prior_J = None

    @property
    def n_intervals(self):
        return self.n_windows * (self.n_frames - 1)

    def host_batch(self, imu_blobs, wheel_blobs, ground_multiplicity=None):
        """The solver input once the streams went through a preintegrator (`imu_blobs` [B(n-1)][466],
        `wheel_blobs` [B(n-1)][15]; ignored when n_frames == 1)."""
        multi = self.n_frames > 1
        return abi.HostBatch(
            self.n_windows, self.n_frames,
            self.ground_multiplicity if ground_multiplicity is None else ground_multiplicity,
            self.prior_frame,
            states=self.states, const_mask=self.const_mask, point_offset=self.point_offset, points=self.points,
            point_line=self.point_line, point_weight=self.point_weight, line_offset=self.line_offset,
            lines=self.lines, ref_frame=self.ref_frame, ref_pose=self.ref_pose,
            imu=imu_blobs if multi else None, wheel=wheel_blobs if multi else None,
            prior_X0=self.prior_X0, prior_J=self.prior_J)


def make_batch(n_windows=1, seed0=42, **kw):
    return SensorBatch([make_window(seed=seed0 + w, **kw) for w in range(n_windows)])


def tile_batch(hb, times):
    """Repeat a HostBatch `times` times (distinct memory, identical contents) to build bench-sized batches
    without ray-casting thousands of windows."""
    a = hb.arrays
    B, n = hb.n_windows, hb.n_frames

    def rep(x):
        return None if x is None else np.tile(x.reshape(B, -1), (times, 1))

    def rep_off(off):
        total = int(off[-1])
        return np.concatenate([(off[:-1][None, :] + total * np.arange(times)[:, None]).ravel(), [total * times]]).astype(np.int64)

    fields = dict(
        states=rep(a["states"]), const_mask=rep(a["const_mask"]),
        point_offset=rep_off(a["point_offset"]), points=np.tile(a["points"].reshape(-1, 2), (times, 1)),
        point_line=np.tile(a["point_line"], times),
        point_weight=None if a["point_weight"] is None else np.tile(a["point_weight"], times),
        line_offset=rep_off(a["line_offset"]), lines=np.tile(a["lines"].reshape(-1, 4), (times, 1)),
        ref_frame=rep(a["ref_frame"]), ref_pose=rep(a["ref_pose"]), imu=rep(a["imu"]), wheel=rep(a["wheel"]),
        prior_X0=rep(a["prior_X0"]), prior_J=rep(a["prior_J"]))
    return abi.HostBatch(B * times, n, hb.ground_multiplicity, hb.prior_frame, **fields)


# ----------------------------------------------------------------------------- the BASELINE.json configs
def config_c1(seed=42):
    """C1: single 361-beam scan vs a 100-segment map, laser factor only, 1 Gauss-Newton/LM iteration."""
    sb = make_batch(1, seed, n_frames=1, beams=361, fov_deg=360.0, topology="tracking", with_prior=False)
    sb.const_mask[:] = 0
    sb.states[0, 0:6] = sb.truth[0, 0:6] + np.array([0.02, -0.015, 0.001, 0.004, -0.003, 0.006])
    sb.ground_multiplicity = 0
    return sb


def config_c2(n_windows=1, seed=42):
    """C2: 1081 beams x 30 frames, fixed associations, IMU + wheel + ground + prior, 10 LM iterations."""
    return make_batch(n_windows, seed, n_frames=30, beams=1081, fov_deg=270.0, topology="tracking")


def config_c4(n_windows=1, seed=42):
    """C4: 4096 beams x 50 frames (synthetic 360 deg scan), points sharded over ranks."""
    return make_batch(n_windows, seed, n_frames=50, beams=4096, fov_deg=360.0, topology="tracking")


def config_init(n_windows=1, seed=42, n_frames=10, mode="segment"):
    """The initialisation window of do_init_solve: slide_window_size = 10 (config/corridor.yaml:73)."""
    return make_batch(n_windows, seed, n_frames=n_frames, beams=1081, fov_deg=270.0, topology="init", mode=mode)


def config_tracking2(n_windows=1, seed=42):
    """The steady-state tracking problem of the reference: 2 frames, matched segment pairs, frame 0 pose constant,
    prior on frame n-2 = 0.  solver::solve adds laser factors for the newest frame only (solver.cpp:669): frame 0's
    pairs hang between two constant poses, so the solver program drops them by itself, while
    solver::marginalization linearises them too (solver.cpp:448-478)."""
    return make_batch(n_windows, seed, n_frames=2, beams=1081, fov_deg=270.0, topology="tracking", mode="segment")


# ----------------------------------------------------------------------------- raw scans for the laser front-end
def make_scan_batch(n_scans=8, seed=42, beams=1081, fov_deg=270.0, n_interior=8, range_sigma=0.01, max_range=30.0,
                    width=10.0, height=8.0, dropout=0.01):
    """Synthetic scans for lvio2d_extract_lines: a width x height room with `n_interior` walls of 1-4 m, the sensor
    at a random pose at least 0.5 m from the outer walls, one ray per beam (-fov/2 .. fov/2), range noise
    N(0, range_sigma), beams without a hit (or randomly dropped with probability `dropout`, like the NaN / inf /
    < 0.1 m readings convert::laser_to_point_times discards, src/utilies/common.cpp:20) removed.
    Returns (point_offset [S+1] int64, points [N][2])."""
    rng = np.random.Generator(np.random.MT19937(int(seed)))
    ang = np.deg2rad(np.linspace(-fov_deg / 2.0, fov_deg / 2.0, beams))
    pts_all, off = [], [0]
    for _ in range(n_scans):
        segs = [[0, 0, width, 0], [width, 0, width, height], [width, height, 0, height], [0, height, 0, 0]]
        for _k in range(n_interior):
            c = rng.uniform([1.0, 1.0], [width - 1.0, height - 1.0])
            a = rng.uniform(0.0, math.pi)
            ln = rng.uniform(1.0, 4.0)
            d = 0.5 * ln * np.array([math.cos(a), math.sin(a)])
            segs.append([*(c - d), *(c + d)])
        world = np.array(segs)
        o = rng.uniform([0.5, 0.5], [width - 0.5, height - 0.5])
        yaw = rng.uniform(-math.pi, math.pi)
        d = np.stack([np.cos(ang + yaw), np.sin(ang + yaw)], axis=1)
        s1, e = world[:, 0:2], world[:, 2:4] - world[:, 0:2]
        so = s1 - o
        den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            rho = (so[None, :, 0] * e[None, :, 1] - so[None, :, 1] * e[None, :, 0]) / den
            u = (so[None, :, 0] * d[:, None, 1] - so[None, :, 1] * d[:, None, 0]) / den
        ok = (np.abs(den) > 1e-12) & (rho > 0.1) & (u >= 0.0) & (u <= 1.0)
        rho = np.where(ok, rho, np.inf)
        r = rho.min(axis=1)
        valid = np.isfinite(r) & (r < max_range) & (rng.uniform(size=beams) >= dropout)
        r = r + rng.normal(0.0, range_sigma, size=beams)
        pts = np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1)[valid]
        pts_all.append(pts)
        off.append(off[-1] + len(pts))
    return np.array(off, dtype=np.int64), np.concatenate(pts_all).reshape(-1, 2)


def make_range_batch(n_scans=8, seed=42, beams=1081, fov_deg=270.0, n_interior=8, range_sigma=0.01, max_range=30.0,
                     width=10.0, height=8.0, scan_period=0.025, moving=True):
    """Raw sensor_msgs/LaserScan-style input for lvio2d_scan_to_points: float32 ranges [S][beams] (inf where no wall is
    hit, a few NaN and sub-0.1 m readings sprinkled in, all of which convert::laser_to_point_times drops) and one
    lvio2d_scan_header per scan (float32 angle_min / angle_increment / time_increment, double stamp, the laser-frame
    velocity used by sensor::laser::correct).  Same room generator as make_scan_batch."""
    rng = np.random.Generator(np.random.MT19937(int(seed)))
    a0 = np.float32(math.radians(-fov_deg / 2.0))
    da = np.float32(math.radians(fov_deg) / (beams - 1))
    ang = (a0 + np.arange(beams, dtype=np.float32) * da).astype(np.float64)
    ranges = np.full((n_scans, beams), np.inf, dtype=np.float32)
    hdr = np.zeros(n_scans, dtype=abi.SCAN_HEADER_DTYPE)
    for k in range(n_scans):
        segs = [[0, 0, width, 0], [width, 0, width, height], [width, height, 0, height], [0, height, 0, 0]]
        for _k in range(n_interior):
            c = rng.uniform([1.0, 1.0], [width - 1.0, height - 1.0])
            a = rng.uniform(0.0, math.pi)
            ln = rng.uniform(1.0, 4.0)
            d = 0.5 * ln * np.array([math.cos(a), math.sin(a)])
            segs.append([*(c - d), *(c + d)])
        world = np.array(segs)
        o = rng.uniform([0.3, 0.3], [width - 0.3, height - 0.3])
        yaw = rng.uniform(-math.pi, math.pi)
        d = np.stack([np.cos(ang + yaw), np.sin(ang + yaw)], axis=1)
        s1, e = world[:, 0:2], world[:, 2:4] - world[:, 0:2]
        so = s1 - o
        den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            rho = (so[None, :, 0] * e[None, :, 1] - so[None, :, 1] * e[None, :, 0]) / den
            u = (so[None, :, 0] * d[:, None, 1] - so[None, :, 1] * d[:, None, 0]) / den
        ok = (np.abs(den) > 1e-12) & (rho > 0.02) & (u >= 0.0) & (u <= 1.0)
        r = np.where(ok, rho, np.inf).min(axis=1)
        r = np.where(np.isfinite(r) & (r < max_range), r + rng.normal(0.0, range_sigma, size=beams), np.inf)
        r[rng.uniform(size=beams) < 0.005] = np.nan
        r[rng.uniform(size=beams) < 0.005] = 0.05
        ranges[k] = r.astype(np.float32)
        hdr[k]["angle_min"], hdr[k]["angle_increment"] = a0, da
        hdr[k]["time_increment"] = np.float32(scan_period / beams)
        hdr[k]["stamp"] = 1.6e9 + 0.1 * k + rng.uniform(0, 0.01)
        if moving:
            hdr[k]["linear"] = rng.normal(0.0, [0.5, 0.05, 0.005])
            hdr[k]["angular"] = rng.normal(0.0, [0.01, 0.01, 0.5])
    return ranges, hdr
