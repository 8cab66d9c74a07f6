#!/bin/bash
# GPU box: the parity tests (all but the large configs unless FULL=1), then per-stage timings of configs 2, 1, 3 (FAST) with the
# in-tree build.  Usage: bash scripts/ab_quick.sh <tag> [configs...]
tag=${1:-ab}; shift
cfgs=${@:-config2 config1 config3}
if [ -n "$FULL" ]; then sel=""; else sel='-k not(config3orconfig4orconfig5)'; fi
if [ -n "$FULL" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/gputest_${tag}.log 2>&1
else python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not config5" > gpurun_out/gputest_${tag}.log 2>&1; fi
tail -3 gpurun_out/gputest_${tag}.log
AB_MODE=1 python scripts/ab_trace.py $tag $cfgs 2>&1 | grep '^{' | tee gpurun_out/ab_${tag}.jsonl
