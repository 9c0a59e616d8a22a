#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_12_pytest.log
tail -15 gpurun_out/r2_12_pytest.log | cut -c1-200
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in old noiseonly powonly default; do ab $lib; done
for lib in old noiseonly powonly default; do ab $lib --camera C3; done
for lib in old default; do ab $lib --camera C1; done
for lib in old default; do ab $lib --width 1920 --height 1080; done
} 2>&1 | tee gpurun_out/r2_12_ab.log
unset RRT_B200_LIB
timeout 300 python bench.py --workload path --steps 1 --path-frames 96 2>&1 | tail -1 | cut -c1-400
