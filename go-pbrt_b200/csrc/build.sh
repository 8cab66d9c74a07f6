#!/bin/bash
# builds libgopbrt_cuda.so in-tree for sm_100a.  -fmad=false / -ffp-contract=off: the reference (Go on amd64) never
# fuses multiply-add, and hit/miss parity is bit-exact.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-ffp-contract=off,-O2,-pthread "$@" -shared -o libgopbrt_cuda.so gopbrt.cu
