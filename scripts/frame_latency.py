"""Device time of every step of the per-frame front-end for ONE robot (M = 1) and for 64: ranges -> points -> lines ->
match_with_ref -> add_scan, all buffers device-resident (on_device = 1), CUDA events on the context's stream."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch
import lvio2d_b200 as L
from lvio2d_b200.solver import Context

lp = L.corridor_line_params()
dev = torch.device("cuda:0")
for M in (1, 64):
    rg1, hd1 = L.synth.make_range_batch(64, 21)
    rg, hd = rg1[:M].copy(), hd1[:M].copy()
    S, nb = rg.shape
    ML = 160
    with Context(L.corridor_params()) as c:
        st = torch.cuda.ExternalStream(c.stream, device=dev)
        d_rg = torch.from_numpy(rg).to(dev)
        d_hd = torch.from_numpy(hd.view(np.uint8).reshape(S, -1).copy()).to(dev)
        d_cnt = torch.zeros(S, dtype=torch.int32, device=dev)
        d_pts = torch.zeros(S * nb * 2, dtype=torch.float64, device=dev)
        d_z = torch.zeros(S * nb, dtype=torch.float64, device=dev)
        d_off = (torch.arange(S, dtype=torch.int64, device=dev) * nb).contiguous()
        d_n = torch.zeros(S, dtype=torch.int32, device=dev)
        d_lines = torch.zeros(S * ML * 4, dtype=torch.float64, device=dev)
        d_abc = torch.zeros(S * ML * 3, dtype=torch.float64, device=dev)
        d_rng = torch.zeros(S * ML * 2, dtype=torch.int32, device=dev)
        d_nm = torch.zeros(S, dtype=torch.int32, device=dev)
        d_m = torch.zeros(S * ML * 2, dtype=torch.int32, device=dev)
        sm = c.submap(lp, S, 4096, 0.01, 0.01, 100)
        poses = np.zeros((S, 6))
        steps = {
            "scan_to_points": lambda dp: c.scan_to_points_device(S, nb, d_rg.data_ptr(), d_hd.data_ptr(), True, d_cnt.data_ptr(), d_pts.data_ptr(), d_z.data_ptr()),
            "extract_lines": lambda dp: c.extract_lines_device(lp, S, d_off.data_ptr(), d_pts.data_ptr(), ML, d_n.data_ptr(), d_lines.data_ptr(), d_abc.data_ptr(),
                                                               d_rng.data_ptr(), point_count_ptr=d_cnt.data_ptr(), point_z_ptr=d_z.data_ptr()),
            "submap_match": lambda dp: sm.match_device(ML, d_n.data_ptr(), d_lines.data_ptr(), dp.data_ptr(), d_nm.data_ptr(), d_m.data_ptr()),
            "submap_add_scan": lambda dp: sm.add_scan_device(ML, d_n.data_ptr(), d_lines.data_ptr(), dp.data_ptr()),
        }
        acc = {k: 0.0 for k in steps}
        warm, reps = 25, 10
        for k in range(warm + reps):
            poses[:, 0], poses[:, 5] = 0.02 * k, 0.003 * k
            dp = torch.from_numpy(poses.reshape(-1).copy()).to(dev)
            torch.cuda.synchronize(dev)
            for name, fn in steps.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); fn(dp); e1.record(st); c.sync()
                if k >= warm:
                    acc[name] += e0.elapsed_time(e1) / reps
        lines_in_map = float(sm.get(0, want_lines=False)[2].mean())
        print(f"M = {M}: " + ", ".join(f"{k} {v * 1e3:.0f} us" for k, v in acc.items()) + f"; sum {sum(acc.values()) * 1e3:.0f} us; sub-map {lines_in_map:.0f} lines")
        sm.close()
