#!/bin/bash
mkdir -p gpurun_out
RRT_PIPELINE=split timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/r2_31_split128 python tools/render_once.py --width 128 --height 72 > gpurun_out/r2_31_ncu.log 2>&1
RRT_KERNEL=scalar timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/r2_31_geo128 python tools/render_once.py --width 128 --height 72 --flags 0 >> gpurun_out/r2_31_ncu.log 2>&1
ls -la gpurun_out/r2_31_*.ncu-rep
