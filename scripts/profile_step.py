"""One device-resident solve of the bench workload (B C2 windows) and nothing else: the target of the ncu captures.
  ncu --set full --clock-control none --import-source on -k regex:factor_pair -s 3 -c 1 -o gpurun_out/x python scripts/profile_step.py
Prints the per-kernel CUDA-event times of the solve."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import lvio2d_b200 as L  # noqa: E402
from lvio2d_b200.solver import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=4096)
ap.add_argument("--solves", type=int, default=2)
ap.add_argument("--config", default="c2")
ap.add_argument("--assoc", default="fixed", choices=["fixed", "nearest"])
args = ap.parse_args()
device = torch.device("cuda", 0)
P = L.corridor_params(max_iters=bench.MAX_ITERS)
P.assoc_mode = 1 if args.assoc == "nearest" else 0
ctx = Context(P)
hb, _ = bench.build_host_batch(ctx, args.windows, seed0=42, config=args.config)
dstruct, keep = bench.to_device_struct(hb, torch, device)
ctx.bind_windows(dstruct, keepalive=keep)
ctx.solve_async()
ctx.sync()
ctx.set_profiling(True)
for _ in range(args.solves):
    ctx.solve_async()
ctx.sync()
prof = ctx.get_profile()
ctx.set_profiling(False)
summ = ctx.get_summaries()
out = {k: (v / max(1, prof[k.replace("_ms", "_launches")]) if k.endswith("_ms") else v) for k, v in prof.items()}
out["iterations_per_window"] = float(summ["iterations"].mean())
print(json.dumps(out))
