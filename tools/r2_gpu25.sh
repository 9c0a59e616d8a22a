#!/bin/bash
# Round 2: ncu source-level capture of trace_kernel (split v2) at 1080p: what do the warps of the slow centre tiles wait for?
mkdir -p gpurun_out
export RRT_PIPELINE=split
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/r2_25_trace1080 python tools/render_once.py --width 1920 --height 1080 > gpurun_out/r2_25_ncu.log 2>&1
tail -2 gpurun_out/r2_25_ncu.log; ls -la gpurun_out/r2_25_trace1080.ncu-rep
