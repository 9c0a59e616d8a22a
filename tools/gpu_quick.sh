#!/bin/bash
# quick A/B: tools/gpu_quick.sh <tag> "<variant> <cam> <extra args>" ...
tag=$1; shift
for spec in "$@"; do
  set -- $spec; v=$1; cam=$2; shift 2
  RRT_B200_LIB=$PWD/build/variants/librrt_b200_variants.so RRT_KERNEL_VARIANT=$v timeout 300 python bench.py --strict --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --camera $cam "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant $v cam $cam $*', 'ms=%.2f'%d['ms_per_step'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'e2e_ms=%.2f'%d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/${tag}_ab.log
