"""Sequence replay through the `lvio_2d::solver` surface — the stand-in for BASELINE.json configs[4].

The OpenLORIS corridor bag is not on disk (no network, no rosbag reader) and the reference's front-end that would
turn it into `laser_match` objects is out of scope, so the replay runs on a synthetic corridor
sequence: one long trajectory through the 100-segment world of `synth.make_window`, matched segment pairs per scan
(the reference's `laser_factor` input), 200 Hz IMU and wheel odometry.  What IS reproduced is the per-frame call
sequence of `trajectory::do_tracking` (reference src/trajectory/trajectory.cpp:523-561):

    opt_solver.solve(frame_infos)  ->  opt_solver.marginalization(frame_infos)  ->  pop_frame_for_tracking()

with a 2-frame window [last laser frame, new frame] (`pop_frame_for_tracking` keeps only the newest laser frame,
trajectory.cpp:592-615), the prior produced by one frame's marginalisation constraining the next frame's solve
(solver.cpp:744-785), and the fast_mode variant (10 iterations, no marginalisation, biases constant,
solver.cpp:258, :790-801).

`run_tracking` is backend-agnostic: it drives any object with the `Solver` methods.  Tests run it once on the CUDA
library and once on the CPU oracle and compare the two trajectories frame by frame.
"""
import numpy as np

from . import synth
from .solver import FrameInfo, LaserMatch, Line


def make_sequence(seed=42, n_frames=40, beams=1081, params=None, **kw):
    """A synthetic corridor sequence in the reference's segment-pair form.  Returns the SensorBatch (one "window" of
    `n_frames` consecutive frames: frame k's scan is matched against the world map expressed in the laser frame of
    the key pose `ref_pose[k]`, as `laser_manager::do_match` would deliver it)."""
    return synth.make_batch(1, seed, n_frames=n_frames, beams=beams, fov_deg=270.0, topology="tracking",
                            mode="segment", params=params, **kw)


def frames_of(sb, imu_blobs, wheel_blobs, frame_dt=0.1):
    """FrameInfo list of a sequence: initial guesses = the batch's perturbed states (stand-in for the IMU
    propagation of trajectory::add_laser_frame), blobs of interval (k-1, k) on frame k."""
    imu = None if imu_blobs is None else np.asarray(imu_blobs).reshape(-1, 466)
    wheel = None if wheel_blobs is None else np.asarray(wheel_blobs).reshape(-1, 15)
    frames = []
    for k in range(sb.n_frames):
        s = sb.states[k]
        f = FrameInfo(frame_dt * k, s[0:3], s[3:6], s[6:9], s[9:15],
                      imu[k - 1] if k else None, wheel[k - 1] if k else None)
        a, b = int(sb.point_offset[k]), int(sb.point_offset[k + 1])
        l0 = int(sb.line_offset[k])
        lines1, lines2 = [], []
        for j in range(a, b, 2):
            li = sb.lines[l0 + sb.point_line[j]]
            lines1.append(Line([li[0], li[1], 0.0], [li[2], li[3], 0.0]))
            lines2.append(Line([*sb.points[j], 0.0], [*sb.points[j + 1], 0.0]))
        f.add_laser_match(LaserMatch(lines1, lines2, sb.ref_pose[k, 0:3], sb.ref_pose[k, 3:6]))
        frames.append(f)
    return frames


def run_tracking(solver, frames, carry_increment=True):
    """trajectory::do_tracking over a sequence.  `frames[0]` is the already-initialised first frame; every later
    frame is appended to the window, solved, marginalised and the older frame popped.  With `carry_increment` the
    correction the solver applied to frame k is carried to the initial guess of frame k+1 (the reference propagates
    the new frame from the last solved state, trajectory.cpp:300-330) so that errors accumulate the way they do in
    a live run.  Returns the solved states [n][15] and the per-frame solver summaries."""
    window = [frames[0]]
    out = [np.concatenate([frames[0].p, frames[0].q, frames[0].v, frames[0].bs])]
    summaries = []
    for k in range(1, len(frames)):
        f = frames[k]
        guess_p, guess_v = f.p.copy(), f.v.copy()
        window.append(f)
        solver.solve(window)
        summaries.append(solver.last_summary)
        solver.marginalization(window)
        window = window[-1:]                                   # pop_frame_for_tracking
        out.append(np.concatenate([f.p, f.q, f.v, f.bs]))
        if carry_increment and k + 1 < len(frames):
            frames[k + 1].p += f.p - guess_p
            frames[k + 1].v += f.v - guess_v
            frames[k + 1].bs[:] = f.bs
    return np.stack(out), summaries


def run_tracking_lockstep(solver, ref_solver, frames):
    """Per-frame comparison on IDENTICAL inputs: `ref_solver` runs the sequence freely; before every solve its
    window (states, laser matches, blobs) and its prior are copied to `solver`, both solve and marginalise, and the
    results of that one frame are compared.  (A free-running comparison is ill-posed in the reference's default
    50-iteration mode: most solves stop at the iteration cap inside the zig-zag regime of the norm-type wheel /
    ground residuals, and the reference path itself turns a 1e-12 input change into 1e-3 m after 40 frames —
    tests/test_sequence.py measures that.)  Returns per-frame (max |dp|, max |dq|, rel. error of the prior
    information J^T J, relative difference of the final costs, 1 if the reference path hit the iteration cap)."""
    import copy

    window = [frames[0]]
    rows = []
    for k in range(1, len(frames)):
        f = frames[k]
        guess_p, guess_v = f.p.copy(), f.v.copy()
        window.append(f)
        mine = copy.deepcopy(window)
        solver.has_linearized_block = ref_solver.has_linearized_block
        solver.linearized_X = None if ref_solver.linearized_X is None else ref_solver.linearized_X.copy()
        solver.linearized_jacobians = None if ref_solver.linearized_jacobians is None else ref_solver.linearized_jacobians.copy()
        ref_solver.solve(window)
        solver.solve(mine)
        dp = max(np.abs(a.p - b.p).max() for a, b in zip(mine, window))
        dq = max(np.abs(a.q - b.q).max() for a, b in zip(mine, window))
        cost, ref_cost = float(solver.last_summary["final_cost"][0]), float(ref_solver.last_summary["final_cost"][0])
        capped = int(ref_solver.last_summary["termination"][0]) == 0      # the reference path stopped at the iteration cap
        # marginalise both at the reference's solution so that the priors are comparable
        for a, b in zip(mine, window):
            a.p[:], a.q[:], a.v[:], a.bs[:] = b.p, b.q, b.v, b.bs
        ref_solver.marginalization(window)
        solver.marginalization(mine)
        dj = 0.0
        if ref_solver.has_linearized_block:
            A = solver.linearized_jacobians.T @ solver.linearized_jacobians
            Bm = ref_solver.linearized_jacobians.T @ ref_solver.linearized_jacobians
            dj = float(np.abs(A - Bm).max() / np.abs(Bm).max())
        rows.append((float(dp), float(dq), dj, (cost - ref_cost) / max(ref_cost, 1e-300), float(capped)))
        window = window[-1:]
        if k + 1 < len(frames):
            frames[k + 1].p += f.p - guess_p
            frames[k + 1].v += f.v - guess_v
            frames[k + 1].bs[:] = f.bs
    return np.array(rows)


def trajectory_rmse(a, b):
    """(position RMSE [m], rotation-vector RMSE [rad]) between two state arrays [n][15]."""
    dp = a[:, 0:3] - b[:, 0:3]
    dq = a[:, 3:6] - b[:, 3:6]
    return float(np.sqrt((dp * dp).sum(axis=1).mean())), float(np.sqrt((dq * dq).sum(axis=1).mean()))
