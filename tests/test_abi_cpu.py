"""CPU-side checks of the drop-in boundary: the CUDA library loads without a GPU, exports every symbol include/lvio2d.h
declares, its structs match the ctypes mirror, and it refuses to run without an sm_100 device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import lvio2d_b200 as L
from lvio2d_b200 import abi, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lvio2d.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lvio2d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(solver.LIB_PATH):
        subprocess.check_call(["bash", os.path.join(ROOT, "2dliw-slam_b200", "csrc", "build.sh")])
    lib = solver.load_library()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/lvio2d.h but not exported"
    assert sorted(solver.EXPORTS) == names, "solver.EXPORTS must list exactly the header's entry points"


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "lvio2d.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(lvio2d_params), sizeof(lvio2d_window_batch), sizeof(lvio2d_summary));
  printf("%zu %zu %zu %zu\n", offsetof(lvio2d_params, g), offsetof(lvio2d_params, max_iters), offsetof(lvio2d_params, huber_delta), offsetof(lvio2d_params, assoc_max_dist));
  printf("%zu %zu %zu\n", offsetof(lvio2d_window_batch, points), offsetof(lvio2d_window_batch, ground_multiplicity), offsetof(lvio2d_window_batch, prior_J));
  printf("%zu %zu %zu %zu %zu\n", sizeof(lvio2d_line_params), sizeof(lvio2d_scan_header), offsetof(lvio2d_scan_header, stamp),
         offsetof(lvio2d_scan_header, linear), offsetof(lvio2d_scan_header, angular));
  return 0; }'''
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    got = [int(x) for x in out]
    want = [C.sizeof(abi.Params), C.sizeof(abi.WindowBatch), C.sizeof(abi.Summary),
            abi.Params.g.offset, abi.Params.max_iters.offset, abi.Params.huber_delta.offset, abi.Params.assoc_max_dist.offset,
            abi.WindowBatch.points.offset, abi.WindowBatch.ground_multiplicity.offset, abi.WindowBatch.prior_J.offset,
            C.sizeof(abi.LineParams), C.sizeof(abi.ScanHeader), abi.ScanHeader.stamp.offset, abi.ScanHeader.linear.offset,
            abi.ScanHeader.angular.offset]
    assert got == want


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(solver.Lvio2dError) as e:
        solver.Context(L.corridor_params())
    assert e.value.status == abi.ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "2dliw-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "liboracle" not in text and "../oracle" not in text, f
