#!/bin/bash
# Builds the CUDA shared library of the C ABI in-tree (sm_100a only).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared -o liblvio2d.so lvio2d_api.cu "$@"
