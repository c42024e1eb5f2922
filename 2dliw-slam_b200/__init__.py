"""lvio2d-b200: B200-native front-end sliding-window solver for 2DLIW-SLAM (`lvio_2d::solver` hot path).

Import as `lvio2d_b200` (alias package at the repo root).  The compute path is the CUDA library
`csrc/liblvio2d.so` reached through the C ABI of include/lvio2d.h; there is no CPU fallback.
"""
from . import abi, params, synth  # noqa: F401
from .params import corridor_line_params, corridor_params  # noqa: F401
