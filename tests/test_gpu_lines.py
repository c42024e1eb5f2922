"""GPU parity of lvio2d_extract_lines (scan -> line segments) against the CPU oracle of
laser_manager::spawn_scan (oracle/laser_lines.hpp), through the C ABI.  Integer outputs (line count, index ranges)
bit-exact; end points and (a, b, c) within 1e-7 (the device takes the smallest eigenvector of the 3x3 Gram matrix, the
oracle the smallest right singular vector of [x y 1]: same vector, condition-number-squared rounding)."""
import numpy as np
import pytest

import lvio2d_b200 as L

pytestmark = pytest.mark.gpu
TOL = 1e-7


@pytest.fixture(scope="module")
def ctx():
    from lvio2d_b200.solver import Context

    c = Context(L.corridor_params())
    yield c
    c.close()


@pytest.fixture(scope="module")
def lp():
    return L.corridor_line_params()


def compare(got, want):
    n, lines, abc, rng = got
    on, olines, oabc, orng = want
    assert np.array_equal(n, on)
    for s in range(len(n)):
        k = min(int(n[s]), lines.shape[1])
        assert np.array_equal(rng[s, :k], orng[s, :k]), f"scan {s}"
        assert np.abs(lines[s, :k] - olines[s, :k]).max(initial=0.0) < TOL
        assert np.abs(abc[s, :k] - oabc[s, :k]).max(initial=0.0) < TOL


@pytest.mark.parametrize("sigma", [0.002, 0.01, 0.03])
def test_lines_match_oracle(ctx, oracle, lp, sigma):
    off, pts = L.synth.make_scan_batch(32, 5, range_sigma=sigma)
    compare(ctx.extract_lines(lp, off, pts), oracle.extract_lines(lp, off, pts))


def test_golden_scans(ctx, lp):
    import os

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lines_scans.npz"))
    got = ctx.extract_lines(lp, G["point_offset"], G["points"], max_lines=G["lines"].shape[1])
    compare(got, (G["n_lines"], G["lines"], G["abc"], G["index_range"]))


def test_edge_cases(ctx, oracle, lp):
    # empty scans, scans shorter than a line, a single point, ragged sizes, an out-of-grid wall
    t = np.linspace(0, 1, 30)[:, None]
    far = (1 - t) * np.array([60.0, 0.0]) + t * np.array([60.5, 3.0])
    off1, pts1 = L.synth.make_scan_batch(3, 9, beams=200)
    pts = np.concatenate([pts1, [[1.0, 0.0], [1.0, 0.1]], [[2.0, 2.0]], far])
    off = np.concatenate([[0, 0], off1[1:], [off1[-1] + 2, off1[-1] + 3, off1[-1] + 3, off1[-1] + 33]]).astype(np.int64)
    got, want = ctx.extract_lines(lp, off, pts), oracle.extract_lines(lp, off, pts)
    compare(got, want)
    assert got[0][0] == 0 and got[0][-1] == 0 and got[0][-2] == 0
    # more lines than slots: the count is still exact, the first max_lines are stored
    off, pts = L.synth.make_scan_batch(2, 3)
    got, want = ctx.extract_lines(lp, off, pts, max_lines=4), oracle.extract_lines(lp, off, pts, max_lines=4)
    assert got[0].min() > 4
    compare(got, want)
    # the wide field of view of BASELINE config 4 (4096 beams, 360 degrees)
    off, pts = L.synth.make_scan_batch(4, 13, beams=4096, fov_deg=360.0)
    compare(ctx.extract_lines(lp, off, pts, max_lines=512), oracle.extract_lines(lp, off, pts, max_lines=512))


def test_bench_size_batch_properties(ctx, lp):
    """4096 scans (the batch bench.py times): every stored line satisfies the filters it passed."""
    off1, pts1 = L.synth.make_scan_batch(64, 21)
    reps = 64
    n1 = int(off1[-1])
    off = np.concatenate([(off1[:-1][None, :] + n1 * np.arange(reps)[:, None]).ravel(), [n1 * reps]]).astype(np.int64)
    pts = np.tile(pts1, (reps, 1))
    n, lines, abc, rng = ctx.extract_lines(lp, off, pts, max_lines=160)
    assert np.array_equal(n.reshape(reps, -1), np.tile(n[:64], (reps, 1)))          # identical scans -> identical results
    assert np.array_equal(lines.reshape(reps, 64, -1), np.tile(lines[:64].reshape(1, 64, -1), (reps, 1, 1)))
    for s in range(0, 64):
        k = int(n[s])
        d = lines[s, :k, 2:] - lines[s, :k, :2]
        assert np.all(np.linalg.norm(d, axis=1) >= lp.line_min_len)
        assert np.all(rng[s, :k, 1] - rng[s, :k, 0] >= 2) and np.all(rng[s, 1:k, 0] >= rng[s, :k - 1, 1])


# ------------------------------------------------------------------ ranges -> points (lvio2d_scan_to_points)
@pytest.mark.parametrize("deskew", [False, True])
def test_scan_to_points_matches_oracle(ctx, oracle, deskew):
    """Counts and time stamps bit-exact (the float32 angle / time arithmetic is replicated rounding by rounding);
    coordinates to 1e-12 (device sincos vs libm)."""
    rg, hd = L.synth.make_range_batch(48, 23)
    rg[5, :] = np.inf          # a scan without a single return
    rg[6, 100:400] = 0.3       # a long run of close readings: every beam falls under the 1 cm filter chain
    cnt, pts, pz, pt = ctx.scan_to_points(rg, hd, deskew=deskew, want_times=True)
    ocnt, opts, opz, opt = oracle.scan_to_points(rg, hd, deskew=deskew)
    assert np.array_equal(cnt, ocnt) and cnt[5] == 0
    for s in range(len(cnt)):
        k = cnt[s]
        assert np.abs(pts[s, :k] - opts[s, :k]).max(initial=0.0) < 1e-12
        assert np.abs(pz[s, :k] - opz[s, :k]).max(initial=0.0) < 1e-12
        assert np.array_equal(pt[s, :k], opt[s, :k])


def test_device_chain_ranges_to_lines(ctx, oracle, lp):
    """ranges -> de-skewed points -> lines, every buffer device-resident between the two calls, against the oracle chain."""
    import torch

    rg, hd = L.synth.make_range_batch(32, 31)
    S, nb = rg.shape
    dev = torch.device("cuda:0")
    d_rg = torch.from_numpy(rg).to(dev)
    d_hd = torch.from_numpy(hd.view(np.uint8).reshape(S, -1).copy()).to(dev)
    d_cnt = torch.zeros(S, dtype=torch.int32, device=dev)
    d_pts = torch.zeros(S * nb * 2, dtype=torch.float64, device=dev)
    d_z = torch.zeros(S * nb, dtype=torch.float64, device=dev)
    d_off = (torch.arange(S, dtype=torch.int64, device=dev) * nb).contiguous()
    ML = 192
    d_n = torch.zeros(S, dtype=torch.int32, device=dev)
    d_lines = torch.zeros(S * ML * 4, dtype=torch.float64, device=dev)
    d_abc = torch.zeros(S * ML * 3, dtype=torch.float64, device=dev)
    d_rng = torch.zeros(S * ML * 2, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.scan_to_points_device(S, nb, d_rg.data_ptr(), d_hd.data_ptr(), True, d_cnt.data_ptr(), d_pts.data_ptr(), d_z.data_ptr())
    ctx.extract_lines_device(lp, S, d_off.data_ptr(), d_pts.data_ptr(), ML, d_n.data_ptr(), d_lines.data_ptr(), d_abc.data_ptr(),
                             d_rng.data_ptr(), point_count_ptr=d_cnt.data_ptr(), point_z_ptr=d_z.data_ptr())
    ctx.sync()
    ocnt, opts, opz, _ = oracle.scan_to_points(rg, hd, deskew=True)
    want = oracle.extract_lines(lp, np.arange(S) * nb, opts.reshape(-1, 2), max_lines=ML, point_count=ocnt, point_z=opz.reshape(-1))
    got = (d_n.cpu().numpy(), d_lines.cpu().numpy().reshape(S, ML, 4), d_abc.cpu().numpy().reshape(S, ML, 3),
           d_rng.cpu().numpy().reshape(S, ML, 2))
    compare(got, want)
    assert got[0].sum() > 10 * S


# ------------------------------------------------------------------ segment association (lvio2d_match_lines)
def _pairs(oracle, lp, n_pairs, seed, guess_noise=0.0):
    """Scan pairs of one world seen from two nearby poses (synthetic windows of 2 frames), their lines and IMU poses."""
    sb = L.synth.make_batch(n_pairs, seed, n_frames=2, beams=1081, n_segments=12, frame_dt=0.3)
    off = sb.point_offset
    n, lines, abc, rng = oracle.extract_lines(lp, off, sb.points, max_lines=128)
    i1, i2 = np.arange(0, 2 * n_pairs, 2), np.arange(1, 2 * n_pairs, 2)
    g = np.random.default_rng(seed)
    pose1, pose2 = sb.truth[i1, :6].copy(), sb.truth[i2, :6] + g.normal(0.0, guess_noise, (n_pairs, 6))
    return dict(n1=n[i1], l1=lines[i1], r1=rng[i1], n2=n[i2], l2=lines[i2], pose1=pose1, pose2=pose2, off1=off[i1],
                cnt1=np.diff(off)[i1].astype(np.int32), pts=sb.points)


@pytest.mark.parametrize("kk,noise", [(0, 0.0), (0, 0.01), (1, 0.02)])
def test_match_lines_matches_oracle(ctx, oracle, lp, kk, noise):
    """Integer outputs: bit-exact, both rasterisation flavours (a line's own points / 0.05 m samples)."""
    P = L.corridor_params()
    d = _pairs(oracle, lp, 24, 100 + kk, noise)
    for flavour in ("points", "samples"):
        kw = dict(point_offset1=d["off1"], points1=d["pts"], index_range1=d["r1"], point_count1=d["cnt1"]) if flavour == "points" else {}
        nm, m = ctx.match_lines(lp, d["n1"], d["l1"], d["n2"], d["l2"], d["pose1"], d["pose2"], kk=kk, **kw)
        onm, om = oracle.match_lines(P, lp, d["n1"], d["l1"], d["n2"], d["l2"], d["pose1"], d["pose2"], kk=kk, **kw)
        assert np.array_equal(nm, onm), flavour
        for p in range(len(nm)):
            assert np.array_equal(m[p, :nm[p]], om[p, :nm[p]]), (flavour, p)
        assert nm.sum() > 5 * len(nm)


def test_match_lines_edge_cases(ctx, oracle, lp):
    P = L.corridor_params()
    d = _pairs(oracle, lp, 4, 7)
    # a pair without lines in scan 2, one without lines in scan 1, and a pose so far off that no cell neighbourhood is hit
    d["n2"][0] = 0
    d["n1"][1] = 0
    d["pose2"][2, 0:2] += 40.0
    nm, m = ctx.match_lines(lp, d["n1"], d["l1"], d["n2"], d["l2"], d["pose1"], d["pose2"], point_offset1=d["off1"], points1=d["pts"],
                            index_range1=d["r1"], point_count1=d["cnt1"])
    onm, om = oracle.match_lines(P, lp, d["n1"], d["l1"], d["n2"], d["l2"], d["pose1"], d["pose2"], point_offset1=d["off1"],
                                 points1=d["pts"], index_range1=d["r1"], point_count1=d["cnt1"])
    assert np.array_equal(nm, onm) and nm[0] == 0 and nm[1] == 0 and nm[2] == 0 and nm[3] > 0
    assert np.array_equal(m[3, :nm[3]], om[3, :nm[3]])


def test_front_end_chain_feeds_the_solver(oracle, lp):
    """points -> lines -> association -> laser_match -> solver::solve, every step on the device path, against the same
    chain on the CPU oracle: the solved pose of the new frame agrees to 1e-6 (bar 1e-4 m / rad)."""
    from lvio2d_b200.solver import Context, FrameInfo, LaserMatch, Line, Solver

    P = L.corridor_params(max_iters=10)
    sb = L.synth.make_batch(3, 77, n_frames=2, beams=1081, n_segments=12, frame_dt=0.3)
    off = sb.point_offset
    hb = oracle.preintegrate_batch(P, sb)
    imu, wheel = hb["imu"].reshape(-1, 466), hb["wheel"].reshape(-1, 15)

    def run(extract, match, make_solver):
        out = []
        n, lines, _, rng = extract(lp, off, sb.points)
        for w in range(3):
            a, b = 2 * w, 2 * w + 1
            nm, m = match(n[[a]], lines[[a]], n[[b]], lines[[b]], sb.truth[[a], :6], sb.states[[b], :6],
                          point_offset1=off[[a]], points1=sb.points, index_range1=rng[[a]], point_count1=np.diff(off)[[a]].astype(np.int32))
            assert nm[0] >= 3
            l1 = [Line([*lines[a, j, :2], 0.0], [*lines[a, j, 2:], 0.0]) for j, _ in m[0, :nm[0]]]
            l2 = [Line([*lines[b, i, :2], 0.0], [*lines[b, i, 2:], 0.0]) for _, i in m[0, :nm[0]]]
            f0 = FrameInfo(0.0, *np.split(sb.truth[a], [3, 6, 9]))
            sg = sb.states[b]
            f1 = FrameInfo(0.3, sg[0:3], sg[3:6], sg[6:9], sg[9:15], imu[w], wheel[w])
            f1.add_laser_match(LaserMatch(l1, l2, sb.truth[a, 0:3], sb.truth[a, 3:6]))
            sol = make_solver()
            sol.solve([f0, f1])
            out.append(np.r_[f1.p, f1.q])
        return np.array(out)

    with Context(P) as c:
        got = run(lambda *a: c.extract_lines(*a, max_lines=128), lambda *a, **k: c.match_lines(lp, *a, **k),
                  lambda: Solver(P, fast_mode=True, ctx=c))
    want = run(lambda *a: oracle.extract_lines(*a, max_lines=128), lambda *a, **k: oracle.match_lines(P, lp, *a, **k),
               lambda: Solver(P, fast_mode=True, ctx=oracle.OracleContext(P)))
    assert np.abs(got - want).max() < 1e-6
    assert np.abs(got - sb.truth[1::2, :6]).max() < np.abs(sb.states[1::2, :6] - sb.truth[1::2, :6]).max()


def test_front_end_error_paths(ctx, lp):
    """Invalid arguments come back as negative status codes (Lvio2dError), never as a crash or a silent fallback."""
    from lvio2d_b200.solver import Lvio2dError

    off, pts = L.synth.make_scan_batch(2, 1, beams=64)
    with pytest.raises(Lvio2dError):
        ctx.extract_lines(lp, off[::-1].copy(), pts)                          # decreasing offsets
    bad = L.corridor_line_params(laser_resolution=0.0)
    with pytest.raises(Lvio2dError):
        ctx.extract_lines(bad, off, pts)
    n, lines, _, rng = ctx.extract_lines(lp, off, pts)
    pose = np.zeros((2, 6))
    with pytest.raises(Lvio2dError):
        ctx.match_lines(lp, n, lines, n, lines, pose, pose, kk=5)             # neighbourhood larger than the 64-bit cell mask
    with pytest.raises(Lvio2dError):
        ctx.match_lines(bad, n, lines, n, lines, pose, pose)
    # zero scans / zero pairs are no-ops
    n0, _, _, _ = ctx.extract_lines(lp, np.zeros(1, np.int64), np.zeros((0, 2)))
    assert len(n0) == 0
