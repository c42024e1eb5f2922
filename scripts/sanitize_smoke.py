"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
every thread-group shape of window_kernel, both factor kernels, the fused small-batch solve, the three-kernel loop,
the split-phase (sharded) entry points, the three laser front-end kernels and the pose-graph kernels (both paths)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import lvio2d_b200 as L
from lvio2d_b200.solver import Context

CASES = (lambda: L.synth.make_batch(2, 42, n_frames=4, beams=90), lambda: L.synth.config_init(1, n_frames=4),
         lambda: L.synth.config_tracking2(1), lambda: L.synth.config_c1())
for env in ({}, {"LVIO2D_WINDOW_THREADS": "32", "LVIO2D_FUSED_SMALL": "0"}, {"LVIO2D_WINDOW_THREADS": "128", "LVIO2D_FUSED_SMALL": "0"},
            {"LVIO2D_WINDOW_THREADS": "512", "LVIO2D_FUSED_SMALL": "0"}, {"LVIO2D_FACTOR_PAIRED": "0", "LVIO2D_FUSED_SMALL": "0"}):
    for k in ("LVIO2D_WINDOW_THREADS", "LVIO2D_FUSED_SMALL", "LVIO2D_FACTOR_PAIRED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for assoc, huber in ((0, 0.0), (1, 1.5)):
        P = L.corridor_params(max_iters=3); P.assoc_mode = assoc; P.huber_delta = huber
        for mk in CASES:
            sb = mk()
            with Context(P) as c:
                hb = c.preintegrate_batch(sb)
                c.set_windows(hb)
                c.linearize(0); c.linearize(1)
                s = c.solve(); x = c.get_states()
                c.marginalize()
                c.set_point_shard(0, 2); c.solve_begin(); c.eval_laser(); c.lm_step(want_active=True); c.set_point_shard(0, 1)
            print("ok", env, assoc, huber, hb.n_frames, s["final_cost"])
# cyclic-reduction shape on a window long enough for several levels and rounds; wire upload with shared sub-map lines and
# the compact IMU records
from lvio2d_b200 import abi
os.environ["LVIO2D_WINDOW_THREADS"] = "512"
os.environ["LVIO2D_FUSED_SMALL"] = "0"
P = L.corridor_params(max_iters=3)
sb = L.synth.make_batch(2, 9, n_frames=13, beams=120, fov_deg=270.0)
with Context(P) as c:
    hb = c.preintegrate_batch(sb)
    c.set_windows(hb)
    s = c.solve()
    import math
    wire = abi.ScanWire.from_points(hb, 120, np.float32(math.radians(-135.0)), np.float32(math.radians(270.0) / 119))
    wire.imu_compact = abi.ScanWire.compact_imu(hb["imu"])
    per = int(np.diff(hb["line_offset"])[0])
    l3 = hb["lines"].reshape(2, 13, per, 4)
    wire.shared_lines = True
    bare = hb.replace(points=None, point_line=None, point_offset=None, imu=None, lines=l3[:, 0].reshape(-1, 4),
                      line_offset=np.arange(3, dtype=np.int64) * per)
    c.set_windows_wire(bare, wire)
    s2 = c.solve()
print("ok cyclic reduction + wire upload", s["final_cost"], s2["final_cost"])
os.environ.pop("LVIO2D_WINDOW_THREADS", None)
os.environ.pop("LVIO2D_FUSED_SMALL", None)
# laser front-end
lp = L.corridor_line_params()
rg, hd = L.synth.make_range_batch(3, 5, beams=181)
rg[1, :] = np.inf
with Context(L.corridor_params()) as c:
    cnt, pts, pz = c.scan_to_points(rg, hd, deskew=True)
    off = np.arange(3, dtype=np.int64) * rg.shape[1]
    n, lines, abc, rng = c.extract_lines(lp, off, pts.reshape(-1, 2), max_lines=64, point_count=cnt, point_z=pz.reshape(-1))
    pose = np.zeros((3, 6))
    nm, m = c.match_lines(lp, n, lines, n, lines, pose, pose, point_offset1=off, points1=pts.reshape(-1, 2), index_range1=rng, point_count1=cnt)
    nm2, m2 = c.match_lines(lp, n, lines, n, lines, pose, pose, kk=1)
    # device-resident reference sub-map: founding, appends through the transform, hand-over, match
    sm = c.submap(lp, 3, 256, 0.01, 0.01, 4)
    for k in range(7):
        pk = np.zeros((3, 6)); pk[:, 0] = 0.05 * k; pk[:, 5] = 0.02 * k
        sm.add_scan(np.where(np.arange(3) == k % 3, -1, n), lines, pk)
    nm3, m3, l1, rp = sm.match(n, lines, pk, want_lines1=True)
    meta = sm.get(0)[0]
    sm.close()
print("ok front-end", cnt, n, nm, nm2, nm3, meta[:, 2])
# back-end pose graph: plain path and the opt-in partitioned path (pose_graph_segments.cuh)
from test_oracle_pose_graph import edge_noise_J
from test_pose_graph_host import graph_with_loops
truth, init, edges, tfs, ws = graph_with_loops(40, [(30, 4), (12, 25), (39, 20)], seed=44)
for segs, stage in (("0", "0"), ("4", "0"), ("2", "1")):
    os.environ["LVIO2D_PG_SEGMENTS"] = segs
    os.environ["LVIO2D_PG_STAGE"] = stage
    with Context(L.corridor_params(max_iters=4)) as c:
        x, s = c.pose_graph_solve(init, edges, tfs, ws, edge_noise_J(), True, True)
        r, J = c.eval_edge_factor(tfs[0], 1.0, edge_noise_J(), init[0], init[1])
    print("ok pose graph, segments", segs, "stage", stage, s["iterations"][0], s["final_cost"][0])
os.environ.pop("LVIO2D_PG_SEGMENTS", None)
os.environ.pop("LVIO2D_PG_STAGE", None)
