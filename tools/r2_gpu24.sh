#!/bin/bash
# Round 2: is the trace kernel's tail the stores to fresh memory?  256-bit stores (default), two 128-bit halves, and a
# timing-only build whose stores all hit one row per warp.
mkdir -p gpurun_out
for v in default st128 dbg4; do
  if [ $v = default ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/$v/librrt_$v.so; fi
  echo "== $v split 1080p"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --width 1920 --height 1080 --reps 4 2>&1 | tail -2 | cut -c1-50
done
unset RRT_B200_LIB
timeout 300 python -m pytest tests/test_gpu_split.py -x -q -k "equals_fused and not 1080" 2>&1 | tail -2
