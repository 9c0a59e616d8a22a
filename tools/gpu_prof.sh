#!/bin/bash
# ncu --set full capture of one render launch.  Usage: tools/gpu_prof.sh <tag> <variant> <bench args...>
tag=$1; v=$2; shift 2
export RRT_KERNEL_VARIANT=$v
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${tag}_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-300
