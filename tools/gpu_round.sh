#!/bin/bash
# One GPU session: parity tests, smoke, A/B bench of the kernel variants.  Usage: tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
ab() {  # variant camera extra...
  v=$1; cam=$2; shift 2
  RRT_B200_LIB=$PWD/build/variants/librrt_b200_variants.so RRT_KERNEL_VARIANT=$v timeout 300 python bench.py --strict --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --camera $cam "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant $v cam $cam $*', 'ms=%.2f'%d['ms_per_step'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'e2e_ms=%.2f'%d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
}
{
ab 3 C0; ab 1 C0
ab 3 C3; ab 1 C3
ab 3 C1; ab 1 C1
ab 3 C0 --flags 0; ab 1 C0 --flags 0
ab 3 C0 --width 1920 --height 1080; ab 1 C0 --width 1920 --height 1080
} 2>&1 | tee gpurun_out/${tag}_ab.log
