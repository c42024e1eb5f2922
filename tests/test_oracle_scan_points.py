"""The ranges -> points oracle (oracle/scan_points.hpp = convert::laser_to_point_times + sensor::laser::correct,
reference src/utilies/common.cpp:4-40, src/trajectory/sensor.h:51-94) against an independent numpy / scipy
restatement (float32 angle arithmetic with numpy scalars, scipy Rotation for make_tf) and hand-built cases."""
import numpy as np
from scipy.spatial.transform import Rotation

import lvio2d_b200 as L
from lvio2d_b200 import abi


def restate(ranges, h, deskew):
    pts, times = [], []
    for i, r in enumerate(ranges):
        if np.isnan(r) or np.isinf(r) or not (float(r) > 0.1):
            continue
        ang = np.float32(h["angle_min"]) + np.float32(np.float32(i) * np.float32(h["angle_increment"]))
        p = np.array([np.cos(float(ang)) * float(r), np.sin(float(ang)) * float(r), 0.0])
        if pts and np.linalg.norm(p - pts[-1]) < 0.01:
            continue
        pts.append(p)
        times.append(float(h["stamp"]) + float(np.float32(np.float32(i) * np.float32(h["time_increment"]))))
    out = []
    for p, t in zip(pts, times):
        if deskew:
            dt = t - float(h["stamp"])
            p = Rotation.from_rotvec(dt * h["angular"]).as_matrix() @ p + dt * h["linear"]
        out.append(p)
    return np.array(out).reshape(-1, 3), np.array(times)


def test_matches_independent_restatement(oracle):
    rg, hd = L.synth.make_range_batch(6, 17)
    for deskew in (False, True):
        cnt, pts, pz, pt = oracle.scan_to_points(rg, hd, deskew=deskew)
        for s in range(len(rg)):
            want, wt = restate(rg[s], hd[s], deskew)
            assert cnt[s] == len(want)
            got = np.c_[pts[s, :cnt[s]], pz[s, :cnt[s]]]
            assert np.abs(got - want).max() < 1e-12
            assert np.array_equal(pt[s, :cnt[s]], wt)


def test_known_answers(oracle):
    hd = np.zeros(1, dtype=abi.SCAN_HEADER_DTYPE)
    hd[0]["angle_min"], hd[0]["angle_increment"], hd[0]["time_increment"], hd[0]["stamp"] = 0.0, np.float32(np.pi / 2), 0.001, 100.0
    # beam 0 along +x, beam 1 along +y, NaN / inf / too-close readings dropped, beam 5 duplicates beam 4's position
    rg = np.array([[2.0, 3.0, np.nan, np.inf, 0.05, 0.05, 1.0, 1.0]], dtype=np.float32)
    cnt, pts, pz, pt = oracle.scan_to_points(rg, hd, deskew=False)
    assert cnt[0] == 4
    assert np.allclose(pts[0, 0], [2.0, 0.0], atol=1e-12) and np.allclose(pts[0, 1], [0.0, 3.0], atol=1e-6)
    assert np.allclose(pt[0, :4] - 100.0, [0.0, 0.001, 0.006, 0.007], atol=1e-6)
    # the 1 cm filter compares with the last KEPT point: three readings 6 mm apart keep the first and the third
    hd[0]["angle_increment"] = np.float32(0.006)
    cnt, pts, _, _ = oracle.scan_to_points(np.array([[1.0, 1.0, 1.0]], dtype=np.float32), hd, deskew=False)
    assert cnt[0] == 2 and abs(np.arctan2(pts[0, 1, 1], pts[0, 1, 0]) - 0.012) < 1e-6
    # pure translation de-skew: p + dt * v
    hd[0]["linear"] = [1.0, 0.0, 0.5]
    cnt, pts, pz, pt = oracle.scan_to_points(np.array([[1.0, 1.0, 1.0]], dtype=np.float32), hd, deskew=True)
    dt = pt[0, 1] - 100.0
    assert abs(pz[0, 1] - 0.5 * dt) < 1e-15 and abs(pts[0, 0, 0] - 1.0) < 1e-15
