"""One K = 1000 pose-graph solve, for an ncu launch list (per-kernel shares of an LM iteration):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pg_launches.csv python scripts/pg_gpu_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lvio2d_b200 as L  # noqa: E402
from lvio2d_b200.solver import Context  # noqa: E402
from test_oracle_pose_graph import edge_noise_J  # noqa: E402
from test_pose_graph_host import graph_with_loops  # noqa: E402

K, nl = 1000, 8
loops = [(K - 10 - 7 * i, 5 + 11 * i) for i in range(nl)]
truth, init, edges, tfs, ws = graph_with_loops(K, loops, seed=K)
with Context(L.corridor_params(max_iters=4)) as ctx:
    got, s = ctx.pose_graph_solve(init, edges, tfs, ws, edge_noise_J(), True, False)
print(s)
