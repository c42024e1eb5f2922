"""world_size-2 `gloo` test of the multi-GPU exchange: every rank owns the point slice [cnt*r/W, cnt*(r+1)/W) of every
frame (the rule of scan_match_kernel / lvio2d_set_point_shard); summing the per-rank normal equations of the laser
terms with one all_reduce reproduces the unsharded ones.  The CPU oracle stands in for the device kernel here (the
exchange logic is what is under test); the device kernel's sharding is tested in test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_batch(hb, rank, world):
    from lvio2d_b200 import abi

    a = hb.arrays
    po = a["point_offset"]
    pts, pl = a["points"].reshape(-1, 2), a["point_line"]
    keep_p, keep_l, off = [], [], [0]
    for f in range(po.size - 1):
        p0, cnt = int(po[f]), int(po[f + 1] - po[f])
        s0, s1 = p0 + cnt * rank // world, p0 + cnt * (rank + 1) // world
        keep_p.append(pts[s0:s1]); keep_l.append(pl[s0:s1])
        off.append(off[-1] + (s1 - s0))
    return abi.HostBatch(hb.n_windows, hb.n_frames, 0, -1, states=a["states"], const_mask=a["const_mask"],
                         point_offset=np.array(off, np.int64), points=np.concatenate(keep_p).reshape(-1, 2),
                         point_line=np.concatenate(keep_l), point_weight=None, line_offset=a["line_offset"], lines=a["lines"],
                         ref_frame=a["ref_frame"], ref_pose=a["ref_pose"], imu=None, wheel=None, prior_X0=None, prior_J=None)


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import lvio2d_b200 as L
    import oracle_lib as O

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = L.corridor_params(max_iters=10)
    sb = L.synth.make_batch(2, 9, n_frames=4, beams=101)
    hb = O.preintegrate_batch(P, sb)
    full = shard_batch(hb, 0, 1)          # laser terms only, all points
    mine = shard_batch(hb, rank, world)
    H, g, c = O.linearize(P, mine)
    buf = torch.from_numpy(np.concatenate([H.ravel(), g.ravel(), c.ravel()]))
    dist.all_reduce(buf)                   # the ONE exchange of an LM iteration
    Hf, gf, cf = O.linearize(P, full)
    want = np.concatenate([Hf.ravel(), gf.ravel(), cf.ravel()])
    err = float(np.abs(buf.numpy() - want).max() / np.abs(want).max())
    n_mine = mine.n_points
    counts = torch.tensor([n_mine], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        out.put((err, int(counts[0]), full.n_points))
    dist.destroy_process_group()


def test_point_shards_allreduce_to_the_unsharded_blocks(oracle):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err, total, want_total = out.get(timeout=10)
    assert total == want_total, "the rank slices must partition every frame's points"
    assert err < 1e-12
