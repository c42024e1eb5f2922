"""Measured fp64 roofline denominators on this GPU (vector DFMA and tensor DMMA), through the C ABI."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lvio2d_b200 as L
from lvio2d_b200.solver import Context

with Context(L.corridor_params(max_iters=10)) as ctx:
    r = [ctx.measure_fp64_peak() for _ in range(3)]
print(json.dumps(r))
