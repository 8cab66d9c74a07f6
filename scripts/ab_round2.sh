python -m pytest tests -m gpu -x -q -k "not config3 and not config4 and not config5" > gpurun_out/gputest_b.log 2>&1; tail -3 gpurun_out/gputest_b.log; grep "\[tier3\]" gpurun_out/gputest_b.log
run() { tag=$1; lib=$2; shift 2
  if [ -n "$lib" ]; then export GOPBRT_LIB=$PWD/go-pbrt_b200/csrc/variants/lib_$lib.so; else unset GOPBRT_LIB; fi
  env "$@" python scripts/ab_trace.py "$tag" $CFGS 2>&1 | grep '^{'
}
CFGS="config2 config1 config4"
run new_fast "" AB_MODE=1
CFGS="config2"
run new_strict "" AB_MODE=0
run noufp_fast "" AB_MODE=1 GOPBRT_NO_UNIFORM_FP=1
run g6_fast g6 AB_MODE=1
run g8_fast g8 AB_MODE=1
run s5_fast s5 AB_MODE=1
