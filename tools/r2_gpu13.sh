#!/bin/bash
mkdir -p gpurun_out
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in default mb20 mb16 k16; do ab $lib; done
for lib in default mb20 mb16; do ab $lib --camera C3; done
} 2>&1 | tee gpurun_out/r2_13_ab.log
unset RRT_B200_LIB
NO_TIMELINE=1 bash tools/r2_gpu9_prof.sh r2b
