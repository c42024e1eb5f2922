"""Host->device copy bandwidth of this box from pinned memory (what bounds bench.py's e2e arm)."""
import json
import torch

res = {}
for mb in (64, 512, 2048):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[f"h2d_{mb}MB_GBps"] = 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[f"d2h_{mb}MB_GBps"] = 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps(res))
