run() { # tag lib env...
  tag=$1; lib=$2; shift 2
  if [ -n "$lib" ]; then export GOPBRT_LIB=$PWD/go-pbrt_b200/csrc/variants/lib_$lib.so; else unset GOPBRT_LIB; fi
  env "$@" AB_MODE=1 python scripts/ab_trace.py "$tag" $CFGS 2>&1 | grep '^{'
}
CFGS="config4"
run base "" X=1
run base_cap24 "" GOPBRT_STACK_CAP=24
run base_cap16 "" GOPBRT_STACK_CAP=16
run b5_cap24 b5 GOPBRT_STACK_CAP=24
run b6_cap24 b6 GOPBRT_STACK_CAP=24
run b8_cap24 b8 GOPBRT_STACK_CAP=24
run b8_cap16 b8 GOPBRT_STACK_CAP=16
