#!/bin/bash
# Round 2: which part of an emission costs the trace kernel its tail?  Timing-only variants (wrong frames) at 1080p.
mkdir -p gpurun_out
for v in default dbg1 dbg2 dbg3; do
  if [ $v = default ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/$v/librrt_$v.so; fi
  echo "== $v split"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --width 1920 --height 1080 --reps 4 2>&1 | tail -2 | cut -c1-60
done
unset RRT_B200_LIB
echo "== fused"; RRT_PIPELINE=fused timeout 200 python tools/render_once.py --width 1920 --height 1080 --reps 3 2>&1 | tail -1 | cut -c1-60
echo "== geodesic only scalar"; RRT_KERNEL=scalar timeout 200 python tools/render_once.py --width 1920 --height 1080 --flags 0 --reps 3 2>&1 | tail -1 | cut -c1-60
