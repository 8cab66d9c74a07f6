#!/bin/bash
# builds a tuning variant of the library: build_variant.sh <name> [-D...]; the result is variants/lib_<name>.so (GOPBRT_LIB selects it)
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p variants
${NVCC:-/usr/local/cuda/bin/nvcc} -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-ffp-contract=off,-O2,-pthread "$@" -shared -o variants/lib_$name.so gopbrt.cu
