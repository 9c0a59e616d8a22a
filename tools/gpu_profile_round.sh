#!/bin/bash
# Default bench run, then the ncu launch list and one --set full capture of the render kernel.  Usage: <tag>
tag=$1
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1; tail -4 gpurun_out/${tag}_bench.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/${tag}_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
