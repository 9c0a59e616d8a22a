#!/bin/bash
# Round 2, GPU session 3: structured-burst variants A/B, parity tests, float-precision census vs the reference CUDA planes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_3_pytest.log
tail -3 gpurun_out/r2_3_pytest.log
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in k0 k4 default k8u2 k16; do ab $lib; done
for lib in k0 default k16; do ab $lib --flags 0; done
for lib in k0 k4 default k8u2 k16; do ab $lib --camera C3; done
for lib in k0 default k16; do ab $lib --camera C1; done
for lib in k0 default; do ab $lib --strict; done
for lib in k0 default; do ab $lib --width 1920 --height 1080; done
} 2>&1 | tee gpurun_out/r2_3_ab.log
unset RRT_B200_LIB
timeout 900 python tests/tools/refcuda_planes_census.py 960 540 --host > gpurun_out/r2_planes_census_960.jsonl 2> gpurun_out/r2_planes_census_960.err
timeout 600 python tests/tools/refcuda_planes_census.py 1920 1080 > gpurun_out/r2_planes_census_1080.jsonl 2> gpurun_out/r2_planes_census_1080.err
tail -2 gpurun_out/r2_planes_census_960.err
