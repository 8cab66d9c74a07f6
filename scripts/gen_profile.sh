#!/bin/bash
# ncu --set full (with source) of the pass's FIRST raygen launch (every lane starts its sample) of the headline frame
tag=${1:-r02c}
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/plain_${tag}.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:"k_generate" --launch-skip 0 --launch-count 1 -f \
    -o gpurun_out/prof_${tag}_gen python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_gen_${tag}.log 2>&1
tail -2 gpurun_out/ncu_gen_${tag}.log
