"""Where one C2 window's solve spends its time: per-kernel device time of the three-kernel loop at B = 1 (and a few
other small batch sizes), from the context's own CUDA events (lvio2d_set_profiling)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch
import lvio2d_b200 as L
from lvio2d_b200.solver import Context
import bench

P = L.corridor_params(max_iters=10)
dev = torch.device("cuda:0")
with Context(P) as c0:
    hb, _ = bench.build_host_batch(c0, 148, seed0=42, config="c2")
for B in (1, 2, 8, 37, 148):
    one = bench.first_windows(hb, B)
    with Context(P) as c:
        d, keep = bench.to_device_struct(one, torch, dev)
        c.bind_windows(d, keepalive=keep)
        ext = torch.cuda.ExternalStream(c.stream, device=dev)
        for _ in range(3):
            c.solve_async()
        c.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(ext)
        for _ in range(reps):
            c.solve_async()
        e1.record(ext)
        c.sync()
        ms = e0.elapsed_time(e1) / reps
        c.set_profiling(True)
        for _ in range(reps):
            c.solve_async()
        c.sync()
        pr = c.get_profile()
        c.set_profiling(False)
        print(f"B={B}: {ms:.3f} ms/solve unprofiled; per launch: scan {pr['scan_ms']/pr['scan_launches']*1e3:.1f} us, factor {pr['factor_ms']/pr['factor_launches']*1e3:.1f} us, "
              f"window {pr['window_ms']/pr['window_launches']*1e3:.1f} us; launches/solve {pr['kernel_launches']/reps:.0f}")
