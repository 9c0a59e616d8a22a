#!/bin/bash
# Round 2: tile timeline of the split pipeline's trace kernel and of the fused kernel (profiling build), with the count of
# checked-step executions per tile (are the lanes of a slow tile still converged?).
mkdir -p gpurun_out
export RRT_B200_LIB=$PWD/build/timeline/librrt_b200_timeline.so
RRT_PIPELINE=split timeout 300 python tools/tile_timeline.py --width 1920 --height 1080 > gpurun_out/r2_18_timeline_1080_split.txt 2>&1
RRT_PIPELINE=fused timeout 300 python tools/tile_timeline.py --width 1920 --height 1080 > gpurun_out/r2_18_timeline_1080_fused.txt 2>&1
cat gpurun_out/r2_18_timeline_1080_split.txt
cat gpurun_out/r2_18_timeline_1080_fused.txt
