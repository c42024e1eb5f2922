"""The device sub-map row in isolation (ncu target): M laser managers accumulate `warm` scans, then match_with_ref is timed.
  ncu --set full --import-source on -k regex:match_lines -s 20 -c 1 -o gpurun_out/x python scripts/submap_probe.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch
import lvio2d_b200 as L
from lvio2d_b200.solver import Context

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lp = L.corridor_line_params()
dev = torch.device("cuda:0")
rg1, hd1 = L.synth.make_range_batch(64, 21)
rg, hd = np.tile(rg1, (M // 64, 1)), np.tile(hd1, M // 64)
S, nb = rg.shape
ML = 160
with Context(L.corridor_params()) as c:
    cnt, pts, pz = c.scan_to_points(rg, hd, deskew=True)
    off = np.arange(S, dtype=np.int64) * nb
    n, lines, _, _ = c.extract_lines(lp, off, pts.reshape(-1, 2), max_lines=ML, point_count=cnt, point_z=pz.reshape(-1))
    d_n, d_lines = torch.from_numpy(n).to(dev), torch.from_numpy(lines.reshape(-1)).to(dev)
    d_nm = torch.zeros(S, dtype=torch.int32, device=dev)
    d_m = torch.zeros(S * ML * 2, dtype=torch.int32, device=dev)
    sm = c.submap(lp, S, 4096, 0.01, 0.01, 100)
    poses = np.zeros((S, 6))
    stream = torch.cuda.ExternalStream(c.stream, device=dev)
    for k in range(warm + 3):
        poses[:, 0] = 0.02 * k
        poses[:, 5] = 0.003 * k
        d_pose = torch.from_numpy(poses.reshape(-1).copy()).to(dev)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sm.match_device(ML, d_n.data_ptr(), d_lines.data_ptr(), d_pose.data_ptr(), d_nm.data_ptr(), d_m.data_ptr())
        e1.record(stream)
        sm.add_scan_device(ML, d_n.data_ptr(), d_lines.data_ptr(), d_pose.data_ptr())
        c.sync()
        if k >= warm:
            print(f"scan {k}: match {e0.elapsed_time(e1):.3f} ms for {S} managers, sub-map lines {sm.get(0, want_lines=False)[2].mean():.0f}, pairs {int(d_nm.sum())}")
    sm.close()
