"""Where does a chunk of the end-to-end arm spend its time?  upload only / upload + solve, with and without imu_compact."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench, lvio2d_b200 as L
from lvio2d_b200 import abi
from lvio2d_b200.solver import Context
P = L.corridor_params(max_iters=10)
ctx = Context(P)
hb, _ = bench.build_host_batch(ctx, 592, seed0=42)
wire = abi.ScanWire.from_points(hb, 1081, np.float32(math.radians(-135.0)), np.float32(math.radians(270.0) / 1080))
wp, wl, _ = wire.points()
hb = hb.replace(points=wp, point_line=wl)
for compact in (False, True, False, True):
    wire.imu_compact = abi.ScanWire.compact_imu(hb["imu"]) if compact else None
    bare = bench.pinned_copy(hb.replace(points=None, point_line=None, point_offset=None, **({"imu": None} if compact else {})), torch)
    wk = bench.pinned_wire(wire, 0, 592, 30, torch)
    for what in ("upload", "upload+solve"):
        ts = []
        for rep in range(6):
            ctx.sync(); t0 = time.perf_counter()
            ctx.set_windows_wire(bare, wk, async_=True)
            t1 = time.perf_counter()
            if what != "upload": ctx.solve_async()
            ctx.sync(); t2 = time.perf_counter()
            ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
        print("compact", compact, what, "host enqueue ms %.2f  total ms %.2f" % tuple(np.median(np.array(ts[2:]), axis=0)), "bytes", bare.nbytes() + wk.nbytes())

# ---- the bench's pipelined loop: 8 chunks over k contexts
nf = 30
out = torch.empty(592 * 8 * nf * 15, dtype=torch.float64).pin_memory().numpy().reshape(-1, 15)
for nctx in (2, 4):
    ctxs = [ctx] + [Context(P) for _ in range(nctx - 1)]
    for compact in (False, True):
        wire.imu_compact = abi.ScanWire.compact_imu(hb["imu"]) if compact else None
        chunks = []
        for k in range(8):
            bare = bench.pinned_copy(hb.replace(points=None, point_line=None, point_offset=None, **({"imu": None} if compact else {})), torch)
            chunks.append((bare, bench.pinned_wire(wire, 0, 592, 30, torch)))
        def step():
            for k, (bare, wk) in enumerate(chunks):
                c = ctxs[k % nctx]
                c.set_windows_wire(bare, wk, async_=True)
                c.solve_async()
                c.get_states_async(out[k * 592 * nf:(k + 1) * 592 * nf])
            for c in ctxs:
                c.sync()
        step(); step()
        t0 = time.perf_counter()
        for _ in range(4): step()
        print("contexts", nctx, "compact", compact, "ms per step %.2f" % ((time.perf_counter() - t0) * 250))
        del chunks
