"""lvio2d_set_windows_wire: the compact wire encoding of a beam-mode batch (float32 ranges + uint16 line index per beam).
CPU: the host statement `abi.ScanWire.points()` against the oracle's convert::laser_to_point_times on scans where the
1 cm thinning cannot trigger.  GPU: a batch uploaded in wire form solves to the same states as the same batch uploaded
with the expanded double points."""
import math

import numpy as np
import pytest

import lvio2d_b200 as L
from lvio2d_b200 import abi

BEAMS, FOV = 361, 270.0
A0 = np.float32(math.radians(-FOV / 2.0))
DA = np.float32(math.radians(FOV) / (BEAMS - 1))


def small_batch(oracle, P, n_windows=3, n_frames=5):
    sb = L.synth.make_batch(n_windows, 7, n_frames=n_frames, beams=BEAMS, fov_deg=FOV)
    hb = oracle.preintegrate_batch(P, sb)
    assert np.all(np.diff(hb["point_offset"]) == BEAMS), "synthetic room is closed: every beam returns"
    return hb


def test_wire_points_follow_laser_to_point_times(oracle):
    rng = np.random.default_rng(3)
    ranges = rng.uniform(8.0, 25.0, size=(4, BEAMS)).astype(np.float32)          # >= 8 m: neighbours are > 1 cm apart
    ranges[0, 5], ranges[1, 7], ranges[2, 9] = np.nan, np.inf, 0.05
    hd = np.zeros(4, dtype=abi.SCAN_HEADER_DTYPE)
    hd["angle_min"], hd["angle_increment"] = A0, DA
    wire = abi.ScanWire(ranges, np.tile(np.array([[A0, DA]], np.float32), (4, 1)), np.zeros((4, BEAMS), np.uint16))
    pts, line, off = wire.points()
    cnt, opts, _, _ = oracle.scan_to_points(ranges, hd, deskew=False)
    for k in range(4):
        keep = line[off[k]:off[k + 1]] >= 0
        assert keep.sum() == cnt[k]
        assert np.array_equal(pts[off[k]:off[k + 1]][keep], opts[k, :cnt[k]])
    assert line[5] == -1 and line[BEAMS + 7] == -1 and line[2 * BEAMS + 9] == -1


def test_compact_imu_keeps_what_the_factor_reads(oracle):
    """imu_factor (imu_factor.h:52-86) reads X, J(0..8, 9..14), the upper triangle of sqrt_inverse_P and Dt: a blob rebuilt
    from the 190 compact doubles gives the same residual and Jacobian."""
    P = L.corridor_params()
    hb = small_batch(oracle, P, n_windows=1, n_frames=3)
    blob = hb["imu"].reshape(-1, abi.IMU_BLOB)[0]
    c = abi.ScanWire.compact_imu(blob)[0]
    assert c.size == abi.IMU_COMPACT
    rebuilt = np.zeros(abi.IMU_BLOB)
    rebuilt[0:15], rebuilt[465] = c[0:15], c[189]
    J = np.zeros((15, 15))
    J[0:9, 9:15] = c[15:69].reshape(9, 6)
    S = np.zeros((15, 15))
    S[np.triu_indices(15)] = c[69:189]
    rebuilt[15:240], rebuilt[240:465] = J.ravel(), S.ravel()
    st = hb["states"].reshape(-1, 15)
    r0, J0 = oracle.eval_imu_factor(P, blob, st[0], st[1])
    r1, J1 = oracle.eval_imu_factor(P, rebuilt, st[0], st[1])
    assert np.array_equal(r0, r1) and np.array_equal(J0, J1)


def test_wire_round_trip_of_a_batch(oracle):
    P = L.corridor_params(max_iters=10)
    hb = small_batch(oracle, P)
    wire = abi.ScanWire.from_points(hb, BEAMS, A0, DA)
    pts, line, off = wire.points()
    assert np.array_equal(off, hb["point_offset"])
    keep = line >= 0                      # beams closer than 0.1 m are rejected like convert::laser_to_point_times rejects them
    assert keep.mean() > 0.95 and np.array_equal(line[keep], hb["point_line"][keep])
    assert np.abs(pts - hb["points"].reshape(-1, 2))[keep].max() < 5e-6       # float32 range and beam grid
    assert wire.nbytes() * 3 < hb["points"].nbytes + hb["point_line"].nbytes


@pytest.mark.gpu
def test_gpu_wire_upload_solves_like_the_expanded_batch(oracle):
    from lvio2d_b200.solver import Context

    P = L.corridor_params(max_iters=10)
    hb = small_batch(oracle, P)
    wire = abi.ScanWire.from_points(hb, BEAMS, A0, DA)
    pts, line, off = wire.points()
    expanded = hb.replace(points=pts, point_line=line, point_offset=off)
    bare = hb.replace(points=None, point_line=None, point_offset=None)
    with Context(P) as c:
        c.set_windows(expanded)
        H0, g0, c0 = c.linearize(0)
        s0 = c.solve()
        x0 = c.get_states()
        c.set_windows_wire(bare, wire)
        H1, g1, c1 = c.linearize(0)
        s1 = c.solve()
        x1 = c.get_states()
    assert np.abs(c1 - c0).max() <= 1e-12 * np.abs(c0).max()
    assert np.abs(H1 - H0).max() <= 1e-12 * np.abs(H0).max() and np.abs(g1 - g0).max() <= 1e-11 * np.abs(g0).max()
    assert np.array_equal(s0["iterations"], s1["iterations"])
    assert np.abs(x1 - x0).max() < 1e-10
    want, _ = oracle.solve(P, expanded)
    assert np.abs(x1 - want).max() < 1e-8
    # the IMU preintegrations as the 190 doubles imu_factor reads instead of the 466 of the blob: identical results
    wire.imu_compact = abi.ScanWire.compact_imu(hb["imu"])
    with Context(P) as c:
        c.set_windows_wire(bare.replace(imu=None), wire)
        s3 = c.solve()
        x3 = c.get_states()
    assert np.array_equal(x3, x1) and np.array_equal(s3["iterations"], s1["iterations"])
    wire.imu_compact = None
    # one line list per window (every frame of a window matched against the same sub-map): identical results
    n, lo = hb.n_frames, hb["line_offset"]
    per = np.diff(lo)
    if np.all(per == per[0]):
        l3 = hb["lines"].reshape(hb.n_windows, n, int(per[0]), 4)
        if np.all(l3 == l3[:, :1]):
            wire.shared_lines = True
            sh = bare.replace(lines=l3[:, 0].reshape(-1, 4), line_offset=np.arange(hb.n_windows + 1, dtype=np.int64) * int(per[0]))
            with Context(P) as c:
                c.set_windows_wire(sh, wire)
                s4 = c.solve()
                x4 = c.get_states()
            assert np.array_equal(x4, x1) and np.array_equal(s4["iterations"], s1["iterations"])
            wire.shared_lines = False
        else:
            pytest.fail("the synthetic windows are expected to share their sub-map lines")
    # 8-bit line indices (local maps of at most 255 lines): identical results
    assert wire.narrow() and wire.nbytes() < 6 * wire.ranges.size
    with Context(P) as c:
        c.set_windows_wire(bare, wire)
        s5 = c.solve()
        x5 = c.get_states()
    assert np.array_equal(x5, x1) and np.array_equal(s5["iterations"], s1["iterations"])
    wire.beam_line8 = None
    # beams without a line take no part: knock out every third beam on the wire and in the expanded batch alike
    wire.beam_line[:, ::3] = abi.ScanWire.NONE
    pts, line, off = wire.points()
    with Context(P) as c:
        c.set_windows_wire(bare, wire)
        c.solve()
        x2 = c.get_states()
    want2, _ = oracle.solve(P, hb.replace(points=pts, point_line=line, point_offset=off))
    assert np.abs(x2 - want2).max() < 1e-8
