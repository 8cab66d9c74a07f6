#!/bin/bash
# ncu --set full (with source) of ONE launch of the Lambert-triangle shade class (wavefront iteration 1) of the headline frame
tag=${1:-r02d}
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/plain_${tag}.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_shade --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/prof_${tag}_shade0 python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${tag}.log
