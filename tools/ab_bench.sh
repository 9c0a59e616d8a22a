#!/bin/bash
# A/B: run the bench on several library builds; prints value / ms per variant
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/$lib; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', 'ms=%.2f'%d['ms_per_step'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
