"""ms per 10-iteration solve against the batch size for every thread-group shape of window_kernel (LVIO2D_WINDOW_THREADS),
C2 (30 frames) and C4 (50 frames): where the cyclic-reduction shape (512) stops paying."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import lvio2d_b200 as L
from lvio2d_b200.solver import Context
import bench

P = L.corridor_params(max_iters=10)
dev = torch.device("cuda:0")
for config, sizes in (("c2", (1, 148, 296, 444, 592, 1184)), ("c4", (1, 64, 256))):
    with Context(P) as c0:
        hb, _ = bench.build_host_batch(c0, max(sizes), seed0=42, config=config)
    for B in sizes:
        one = bench.first_windows(hb, B)
        row = []
        for wt in (0, 512, 256, 128, 32):
            if wt:
                os.environ["LVIO2D_WINDOW_THREADS"] = str(wt)
            else:
                os.environ.pop("LVIO2D_WINDOW_THREADS", None)
            with Context(P) as c:
                d, keep = bench.to_device_struct(one, torch, dev)
                c.bind_windows(d, keepalive=keep)
                ext = torch.cuda.ExternalStream(c.stream, device=dev)
                for _ in range(3):
                    c.solve_async()
                c.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record(ext)
                for _ in range(reps):
                    c.solve_async()
                e1.record(ext)
                c.sync()
                row.append(f"{wt or 'auto'}: {e0.elapsed_time(e1) / reps:.3f}")
        print(config, "B =", B, "ms/solve ->", "  ".join(row), flush=True)
