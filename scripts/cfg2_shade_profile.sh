#!/bin/bash
# ncu --set full (with source) of the shade + flat-trace kernels of wavefront iteration 1 of the headline frame (config 2, FAST)
tag=${1:-r02b}
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/plain_${tag}.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:"k_shade|k_trace_flat" --launch-skip 6 --launch-count 6 -f \
    -o gpurun_out/prof_${tag}_cfg2 python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
