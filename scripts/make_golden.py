#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU oracle (the reference ships no golden vectors and cannot run here).
Each file freezes the complete solver input (the lvio2d_window_batch arrays) together with the oracle's outputs, so
the fixtures do not depend on the synthetic generator staying bit-stable.  Re-run: python scripts/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lvio2d_b200 as L  # noqa: E402
import oracle_lib as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def pack(hb):
    d = {"n_windows": hb.n_windows, "n_frames": hb.n_frames, "ground_multiplicity": hb.ground_multiplicity, "prior_frame": hb.prior_frame}
    for k, v in hb.arrays.items():
        if v is not None:
            d["in_" + k] = v
    return d


def window_case(name, sb, iters):
    P = L.corridor_params(max_iters=iters)
    hb = O.preintegrate_batch(P, sb)
    d = pack(hb)
    d["max_iters"] = iters
    for mode in (0, 1):
        H, g, c = O.linearize(P, hb, mode=mode)
        d[f"H{mode}"], d[f"g{mode}"], d[f"cost{mode}"] = H, g, c
    st, summ = O.solve(P, hb)
    d["solved_states"] = st
    for f in summ.dtype.names:
        d["summary_" + f] = summ[f]
    X0, J, r, dH, dg = O.marginalize(P, hb)
    d["marg_X0"], d["marg_J"], d["marg_r"], d["marg_dH"], d["marg_dg"] = X0, J, r, dH, dg
    if sb.n_frames > 1:
        d["raw_imu_offset"], d["raw_imu_samples"], d["raw_bias0"] = sb.imu_offset, sb.imu_samples, sb.bias0
        d["raw_wheel_offset"], d["raw_wheel_steps"] = sb.wheel_offset, sb.wheel_steps
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "cost", d["cost0"], "->", summ["final_cost"], "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def factor_case():
    P = L.corridor_params()
    rng = np.random.default_rng(2024)
    sb = L.synth.make_batch(1, 77, n_frames=3, beams=16)
    hb = O.preintegrate_batch(P, sb)
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)
    s = sb.states
    d = {"imu_blob": imu[0], "wheel_blob": wheel[0], "state_i": s[0], "state_j": s[1]}
    d["imu_res"], d["imu_jac"] = O.eval_imu_factor(P, imu[0], s[0], s[1])
    d["wheel_res"], d["wheel_jac"] = O.eval_wheel_factor(P, wheel[0], s[0, :6], s[1, :6])
    d["ground_res"], d["ground_jac"] = O.eval_ground_factors(P, s[1, :6])
    l = np.concatenate([rng.uniform(-6, 6, (4, 2)), np.zeros((4, 1))], axis=1)
    d["laser_lines"] = l
    d["laser_res"], d["laser_jac"] = O.eval_laser_factor(P, l[0], l[1], l[2], l[3], s[0, :6], s[2, :6])
    d["laser_pose_j"] = s[2, :6]
    np.savez_compressed(os.path.join(OUT, "factors.npz"), **d)
    print("factors ok")


def lines_case():
    """Scans + the oracle's scan::lines (laser_manager::spawn_scan): inputs are frozen with the outputs."""
    lp = L.corridor_line_params()
    off, pts = L.synth.make_scan_batch(6, 2024)
    n, lines, abc, rng = O.extract_lines(lp, off, pts, max_lines=128)
    np.savez_compressed(os.path.join(OUT, "lines_scans.npz"), point_offset=off, points=pts, n_lines=n, lines=lines, abc=abc,
                        index_range=rng)
    print("lines ok", n)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    O.build()
    if "--lines" in sys.argv:
        lines_case()
        sys.exit(0)
    factor_case()
    lines_case()
    window_case("c1_single_scan", L.synth.config_c1(), 1)
    window_case("c2_small", L.synth.make_batch(2, 42, n_frames=5, beams=120, fov_deg=270.0), 10)
    window_case("tracking2_segments", L.synth.config_tracking2(1), 20)
    window_case("init_segments", L.synth.config_init(1, n_frames=5), 20)
