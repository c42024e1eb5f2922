#!/bin/bash
# Build kernel variants of the CUDA library side by side (csrc/variants/lib_<name>.so) for A/B timing on the GPU box:
#   scripts/variants.sh name1 "-DFLAG=1" name2 "-DFLAG=2 -DOTHER" ...
# Select one at run time with LVIO2D_LIB=<path> (2dliw-slam_b200/solver.py).
set -e
cd "$(dirname "$0")/../2dliw-slam_b200/csrc"
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared $flags \
    -o variants/lib_$name.so lvio2d_api.cu 2>/dev/null &
done
wait
ls -la variants
