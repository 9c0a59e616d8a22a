#!/bin/bash
mkdir -p gpurun_out
RRT_KERNEL=packed timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_10_pytest_packed.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_10_pytest_packed.log
tail -25 gpurun_out/r2_10_pytest_packed.log
ab() {  # kernel extra...
  k=$1; shift
  RRT_KERNEL=$k timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$k $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for k in scalar packed; do ab $k; done
for k in scalar packed; do ab $k --flags 0; done
for k in scalar packed; do ab $k --camera C3; done
for k in scalar packed; do ab $k --camera C1; done
for k in scalar packed; do ab $k --width 1920 --height 1080; done
for k in scalar packed; do ab $k --width 256 --height 256 --flags 0; done
} 2>&1 | tee gpurun_out/r2_10_ab.log
