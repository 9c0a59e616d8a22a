#!/bin/bash
# Round 2: ncu source-level capture of media_kernel on the 4K bench frame.
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:media_kernel -s 2 -c 1 -f -o gpurun_out/r2_36_media4k python tools/render_once.py > gpurun_out/r2_36_ncu.log 2>&1
tail -2 gpurun_out/r2_36_ncu.log; ls -la gpurun_out/r2_36_media4k.ncu-rep
