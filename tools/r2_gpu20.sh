#!/bin/bash
# Round 2: where does a lone disk-plane tile of trace_kernel wait?  ncu source-level capture of a 128x72 split frame.
mkdir -p gpurun_out
export RRT_PIPELINE=split
timeout 300 python tools/render_once.py --width 128 --height 72 > gpurun_out/r2_20_plain.log 2>&1; cat gpurun_out/r2_20_plain.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/r2_20_trace128 python tools/render_once.py --width 128 --height 72 > gpurun_out/r2_20_ncu.log 2>&1
tail -3 gpurun_out/r2_20_ncu.log; ls -la gpurun_out/*.ncu-rep
