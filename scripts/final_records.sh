#!/bin/bash
# One GPU box, the in-tree build: every single-GPU record under profiles/ (names: gpurun_out/<tag>_*).  Usage: final_records.sh <tag> [light]
tag=${1:-r02}; light=$2
python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputest.log 2>&1; tail -2 gpurun_out/${tag}_gputest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 400 gpurun_out/${tag}_bench_n1.json
[ -n "$light" ] && exit 0
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/${tag}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launch_list_raw.csv \
    python bench.py --profile --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_trace|k_shade|k_split" --launch-skip 6 --launch-count 6 -f \
    -o gpurun_out/${tag}_prof_cfg2 python bench.py --profile --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_full.log 2>&1
AB_MODE=1 ITER_LOG_NAME=${tag}_iter_config2_fast.csv python scripts/iter_log.py config2 > gpurun_out/${tag}_iter_log.txt 2>&1
python scripts/fullsize_parity.py > gpurun_out/${tag}_fullsize_parity.json 2> gpurun_out/${tag}_fullsize_parity.err
tail -c 600 gpurun_out/${tag}_fullsize_parity.json
