"""the device-resident reference sub-map (lvio2d_submap_*: laser_manager::add_scan's state
machine and match_with_ref, reference src/trajectory/laser_manager.cpp:424-496, :531-546) against
  * the reference's own laser_manager class (oracle/_ref/libref.so, reference text compiled unmodified), frame by frame,
  * the host mirror lvio2d_b200.frontend.LaserManager on the oracle back-end (itself pinned to the reference text by
    tests/test_ref_frontend.py), for a batch of managers with different motion.
Counts, flags and match pairs must be equal; poses bit-equal (they are copies); line end points to 1e-9 m (one rigid
transform of coordinates < 30 m; measured ~1e-14)."""
from types import SimpleNamespace

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import lvio2d_b200 as L
import ref_lib
from lvio2d_b200.frontend import LaserManager

pytestmark = pytest.mark.gpu



@pytest.fixture(scope="module")
def ctx():
    from lvio2d_b200.solver import Context

    c = Context(L.corridor_params())
    yield c
    c.close()


def _trajectory(seed, frames, still_every=5):
    """IMU poses of a robot creeping along a corridor; every `still_every`-th frame repeats the last pose."""
    P = L.corridor_params()
    T_il = np.array(list(P.T_imu_to_laser)).reshape(3, 4)
    base = Rotation.from_matrix(T_il[:, :3].T)
    g = np.random.default_rng(seed)
    xy, yaw, out = np.zeros(2), 0.0, []
    for f in range(frames):
        if f % still_every:
            xy = xy + g.normal(0, 0.02, 2) + np.array([0.015, 0.0])
            yaw += g.normal(0, 0.01)
        Rwi = Rotation.from_euler("z", yaw).as_matrix() @ base.as_matrix()
        out.append(np.r_[xy, 0.0, Rotation.from_matrix(Rwi).as_rotvec()])
    return np.array(out)


def _scans(oracle, lp, n, seed, max_lines):
    off, pts = L.synth.make_scan_batch(n, seed, beams=721, range_sigma=0.004)
    nl, lines, _, _ = oracle.extract_lines(lp, off, pts, max_lines=max_lines)
    return off, pts, nl, lines


def _check_against(sm, m, host, frame):
    """sub-map state of manager m on the device against a host LaserManager-like object."""
    for which, sub in ((0, host.ref_submap_ptr), (1, host.spawnning_ref_submap_ptr)):
        meta, pose, n, lines = sm.get(which)
        assert bool(meta[m, which]) == (sub is not None), (frame, which)
        if sub is None:
            continue
        assert meta[m, 2] == host.current_count, (frame, meta[m], host.current_count)
        assert np.array_equal(pose[m], np.r_[sub.current_p, sub.current_q]), (frame, which)
        want = np.array([[l.p1[0], l.p1[1], l.p2[0], l.p2[1]] for l in sub.scan_ptr.lines]).reshape(-1, 4)
        assert n[m] == len(want), (frame, which, n[m], len(want))
        if len(want):
            assert np.abs(lines[m, :n[m]] - want).max() < 1e-9, (frame, which)


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not built and /root/reference absent")
def test_submap_matches_reference_text(ctx, oracle):
    """One manager, 260 frames (two hand-overs at ref_n_accumulation = 100), against lvio_2d::laser_manager itself."""
    lp = L.corridor_line_params()
    ML = 160
    off, pts, nl, lines = _scans(oracle, lp, 4, 31, ML)
    poses = _trajectory(8, 260)
    ref = ref_lib.RefLaserManager()
    sm = ctx.submap(lp, 1, 16384, 0.01, 0.01, ref_lib.REF_N_ACCUMULATION)
    rolled = matched = 0
    last_ref_pose = None
    for f, pose in enumerate(poses):
        k = f % 4
        p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
        rs = ref_lib.RefScan.from_points(p3)
        assert rs.num_lines() == nl[k]
        if f % 7 == 3:     # trajectory.cpp's order: match against the sub-map as it stands, then add
            want, rpose = ref.match_with_ref(rs, pose)
            nm, m = sm.match(nl[k:k + 1], lines[k:k + 1], pose)
            assert nm[0] == len(want) and np.array_equal(m[0, :nm[0]], want), f
            matched += len(want)
        ref.add_scan(rs, pose)
        sm.add_scan(nl[k:k + 1], lines[k:k + 1], pose)
        for which in (0, 1):
            r = ref.submap(which)
            meta, dpose, n, dl = sm.get(which)
            assert bool(meta[0, which]) == (r is not None), (f, which)
            if r is None:
                continue
            rpose, rlines, rcount = r
            assert meta[0, 2] == rcount and n[0] == len(rlines), (f, which, meta[0], rcount, n[0], len(rlines))
            assert np.array_equal(dpose[0], rpose), (f, which)
            if len(rlines):
                assert np.abs(dl[0, :n[0]] - rlines[:, [0, 1, 3, 4]]).max() < 1e-9, (f, which)
        r0 = ref.submap(0)
        if last_ref_pose is not None and np.abs(r0[0] - last_ref_pose).max() > 0:
            rolled += 1
        last_ref_pose = r0[0].copy()
    assert rolled >= 2 and matched > 100
    sm.close()


def test_submap_batch_matches_host_mirror(ctx, oracle):
    """Three managers with their own trajectories (one of them fed only every other call), a small ref_n_accumulation so
    that the buffers roll several times, against frontend.LaserManager on the oracle back-end; then match_with_ref."""
    P, lp = L.corridor_params(), L.corridor_line_params()
    ML, M, n_acc = 96, 3, 12
    off, pts, nl, lines = _scans(oracle, lp, 6, 17, ML)
    assert nl.max() <= ML
    be = oracle.OracleContext(P)
    hosts = [LaserManager(be, lp, max_lines=ML, params=P, ref_n_accumulation=n_acc) for _ in range(M)]
    sm = ctx.submap(lp, M, 4096, 0.01, 0.01, n_acc)
    traj = [_trajectory(40 + m, 90, still_every=4 + m) for m in range(M)]
    scans = []
    for k in range(6):
        p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
        scans.append(hosts[0].spawn_scan(SimpleNamespace(points=p3, times=np.zeros(1), time_stamp=0.0)))
        assert len(scans[-1].lines) == nl[k]
    matched = 0
    for f in range(90):
        n_in, l_in, pose_in = np.zeros(M, np.int32), np.zeros((M, ML, 4)), np.zeros((M, 6))
        ks = [(f + 2 * m) % 6 for m in range(M)]
        fed = [not (m == 2 and f % 2) for m in range(M)]
        for m in range(M):
            n_in[m] = nl[ks[m]] if fed[m] else -1
            l_in[m] = lines[ks[m]]
            pose_in[m] = traj[m][f]
        if f % 5 == 2:
            nm, mt = sm.match(nl[ks], l_in, pose_in)
            for m in range(M):
                hm = hosts[m].match_with_ref(scans[ks[m]], pose_in[m, 0:3], pose_in[m, 3:6])
                sub = hosts[m].ref_submap_ptr
                want = np.array([[sub.scan_ptr.lines.index(a), scans[ks[m]].lines.index(b)] for a, b in zip(hm.lines1, hm.lines2)],
                                np.int32).reshape(-1, 2) if sub is not None else np.zeros((0, 2), np.int32)
                assert nm[m] == len(want) and np.array_equal(mt[m, :nm[m]], want), (f, m)
                matched += len(want)
        sm.add_scan(n_in, l_in, pose_in)
        for m in range(M):
            if fed[m]:
                hosts[m].add_scan(scans[ks[m]], pose_in[m, 0:3], pose_in[m, 3:6])
            _check_against(sm, m, hosts[m], f)
    meta = sm.get(0, want_lines=False)[0]
    assert (meta[:, 0] == 1).all() and (meta[:, 1] == 1).all() and matched > 200
    # line_cap: lines beyond the capacity are counted but not stored, the rest of the state is untouched
    small = ctx.submap(lp, 1, 8, 0.01, 0.01, n_acc)
    small.add_scan(nl[0:1], lines[0:1], traj[0][0])
    meta, _, n, dl = small.get(0)
    big = ctx.submap(lp, 1, 4096, 0.01, 0.01, n_acc)
    big.add_scan(nl[0:1], lines[0:1], traj[0][0])
    _, _, nb, bl = big.get(0)
    assert meta[0, 0] == 1 and n[0] == nb[0] > 8 and np.array_equal(dl[0], bl[0, :8])
    big.close()
    small.reset()
    assert small.get(0, want_lines=False)[0].sum() == 0
    sm.close()
    small.close()


def test_submap_device_chain(ctx, oracle):
    """extract_lines -> submap_add_scan -> submap_match with every buffer in device memory (on_device = 1), against the
    host-buffer flavour of the same entry points."""
    import torch

    lp = L.corridor_line_params()
    M, ML = 8, 128
    off, pts = L.synth.make_scan_batch(M, 77, beams=721, range_sigma=0.004)
    dev = torch.device("cuda:0")
    d_off, d_pts = torch.from_numpy(off).to(dev), torch.from_numpy(pts.reshape(-1)).to(dev)
    d_n = torch.zeros(M, dtype=torch.int32, device=dev)
    d_lines = torch.zeros(M * ML * 4, dtype=torch.float64, device=dev)
    d_abc = torch.zeros(M * ML * 3, dtype=torch.float64, device=dev)
    d_rng = torch.zeros(M * ML * 2, dtype=torch.int32, device=dev)
    poses = np.array([_trajectory(60 + m, 3)[1:] for m in range(M)])       # [M][2][6]
    d_pose = [torch.from_numpy(poses[:, j].copy()).to(dev) for j in range(2)]
    d_nm = torch.zeros(M, dtype=torch.int32, device=dev)
    d_match = torch.zeros(M * ML * 2, dtype=torch.int32, device=dev)
    a, b = ctx.submap(lp, M, 2048), ctx.submap(lp, M, 2048)
    torch.cuda.synchronize()
    ctx.extract_lines_device(lp, M, d_off.data_ptr(), d_pts.data_ptr(), ML, d_n.data_ptr(), d_lines.data_ptr(), d_abc.data_ptr(), d_rng.data_ptr())
    for j in range(2):
        a.add_scan_device(ML, d_n.data_ptr(), d_lines.data_ptr(), d_pose[j].data_ptr())
    a.match_device(ML, d_n.data_ptr(), d_lines.data_ptr(), d_pose[1].data_ptr(), d_nm.data_ptr(), d_match.data_ptr())
    ctx.sync()
    nl, lines = d_n.cpu().numpy(), d_lines.cpu().numpy().reshape(M, ML, 4)
    for j in range(2):
        b.add_scan(nl, lines, poses[:, j])
    nm, mt = b.match(nl, lines, poses[:, 1])
    for which in (0, 1):
        for x, y in zip(a.get(which), b.get(which)):
            assert np.array_equal(x, y)
    assert np.array_equal(d_nm.cpu().numpy(), nm) and np.array_equal(d_match.cpu().numpy().reshape(M, ML, 2), mt)
    assert nm.sum() > 10 * M and (a.get(0, want_lines=False)[2] > nl).all()
    ptrs = a.device_pointers()
    assert all(ptrs)
    a.close()
    b.close()


def test_frontend_laser_manager_on_device_submap(ctx, oracle):
    """frontend.LaserManager(device_submap=True): the same LaserMatch (pairs, sub-map pose, matched end points) as the
    host bookkeeping."""
    P, lp = L.corridor_params(), L.corridor_line_params()
    off, pts = L.synth.make_scan_batch(4, 5, beams=721, range_sigma=0.004)
    host = LaserManager(ctx, lp, max_lines=160, params=P, ref_n_accumulation=10)
    devm = LaserManager(ctx, lp, max_lines=160, params=P, ref_n_accumulation=10, device_submap=True, line_cap=4096)
    poses = _trajectory(3, 40)
    pairs = 0
    for f, pose in enumerate(poses):
        k = f % 4
        p3 = np.c_[pts[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])]
        sc = host.spawn_scan(SimpleNamespace(points=p3, times=np.zeros(1), time_stamp=0.0))
        mh, md = host.match_with_ref(sc, pose[0:3], pose[3:6]), devm.match_with_ref(sc, pose[0:3], pose[3:6])
        assert len(mh.lines1) == len(md.lines1), f
        assert np.array_equal(mh.p1, md.p1) and np.array_equal(mh.q1, md.q1)
        for x, y, u, v in zip(mh.lines1, md.lines1, mh.lines2, md.lines2):
            assert u is v and np.abs(x.p1 - y.p1).max() < 1e-9 and np.abs(x.p2 - y.p2).max() < 1e-9
        pairs += len(mh.lines1)
        host.add_scan(sc, pose[0:3], pose[3:6])
        devm.add_scan(sc, pose[0:3], pose[3:6])
    assert pairs > 200
