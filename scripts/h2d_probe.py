"""Host->device copy bandwidth of this box from pinned memory (what bounds bench.py's e2e arm): per-copy time as a
function of size, so that the fixed cost of one cudaMemcpyAsync can be separated from the link bandwidth."""
import json
import torch

res = {}
big = torch.empty(2048 * 1024 * 1024, dtype=torch.uint8).pin_memory()
dev = torch.empty_like(big, device="cuda")
for mb in (1, 4, 16, 64, 256, 1024, 2048):
    n = mb * 1024 * 1024
    reps = max(2, min(64, 2048 // mb))
    h, d = big[:n], dev[:n]
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"h2d_{mb}MB"] = {"ms_per_copy": ms, "GBps": n / (ms * 1e-3) / 1e9}
# many arrays of one batch: 13 separate copies vs one copy of the same total (512 MB)
parts = [big[i * 40 * 1024 * 1024:(i + 1) * 40 * 1024 * 1024] for i in range(13)]
dparts = [dev[i * 40 * 1024 * 1024:(i + 1) * 40 * 1024 * 1024] for i in range(13)]
for label, fn in (("13x40MB", lambda: [d.copy_(h, non_blocking=True) for h, d in zip(parts, dparts)]),
                  ("1x520MB", lambda: dev[:520 * 1024 * 1024].copy_(big[:520 * 1024 * 1024], non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[label] = {"ms": e0.elapsed_time(e1) / 4, "GBps": 520 * 1024 * 1024 / (e0.elapsed_time(e1) / 4 * 1e-3) / 1e9}
print(json.dumps(res, indent=1))
