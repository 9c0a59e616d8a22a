#!/bin/bash
# Round 2, GPU session 1: RF-bank microbenchmarks, parity tests with the burst loop, A/B of burst lengths.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_1_gpu.txt
timeout 120 tools/microbench/bin/rf_banks > gpurun_out/r2_rf_banks.txt 2>&1
timeout 300 tools/microbench/bin/step_bench > gpurun_out/r2_step_bench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_1_pytest.log
tail -3 gpurun_out/r2_1_pytest.log
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in k0 k2 default k8; do ab $lib; done
for lib in k0 k2 default k8; do ab $lib --flags 0; done
for lib in k0 default k8; do ab $lib --camera C3; done
for lib in k0 default k8; do ab $lib --strict; done
} 2>&1 | tee gpurun_out/r2_1_ab.log
