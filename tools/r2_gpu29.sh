#!/bin/bash
# Round 2: do the lanes of a slow tile stay converged?  Tile timeline with the count of checked-step executions per tile.
mkdir -p gpurun_out
export RRT_B200_LIB=$PWD/build/gs/librrt_gs.so
for cfg in "split 3" "fused 3" "fused 0"; do
  set -- $cfg
  echo "=== $1 flags=$2 128x72"
  RRT_KERNEL=scalar RRT_PIPELINE=$1 timeout 300 python tools/tile_timeline.py --width 128 --height 72 --flags $2 2>&1 | sed -n '1,3p;12,20p;31,32p'
done
echo "=== split flags=3 1080p"
RRT_PIPELINE=split timeout 300 python tools/tile_timeline.py --width 1920 --height 1080 2>&1 | sed -n '1,3p;12,20p;31,32p'
