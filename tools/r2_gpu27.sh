#!/bin/bash
# Round 2: a 128x72 frame (every tile alone on its SM sub-partition): frame time = the slowest tile's own latency.
mkdir -p gpurun_out
for v in default dbg1 dbg3 dbg4; do
  if [ $v = default ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/$v/librrt_$v.so; fi
  echo "== $v split"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --width 128 --height 72 --reps 4 2>&1 | tail -2 | cut -c1-36
done
unset RRT_B200_LIB
echo "== fused"; RRT_PIPELINE=fused timeout 200 python tools/render_once.py --width 128 --height 72 --reps 3 2>&1 | tail -1 | cut -c1-36
echo "== geodesic only scalar"; RRT_KERNEL=scalar timeout 200 python tools/render_once.py --width 128 --height 72 --flags 0 --reps 3 2>&1 | tail -1 | cut -c1-36
echo "== geodesic only packed"; RRT_KERNEL=packed timeout 200 python tools/render_once.py --width 128 --height 72 --flags 0 --reps 3 2>&1 | tail -1 | cut -c1-36
