#!/bin/bash
# ncu captures of the headline frame (config 2, 1080p, FAST): launch list of one frame + `--set full` of the stage kernels of
# wavefront iterations 1-2.  Run on a GPU box: bash scripts/cfg2_profile.sh <tag>
tag=${1:-r02}
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/plain_${tag}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_list_${tag}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_trace|k_shade|k_generate|k_split" --launch-skip 8 --launch-count 12 -f \
    -o gpurun_out/prof_${tag}_cfg2 python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
