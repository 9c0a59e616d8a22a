#!/bin/bash
# Round 2 profile session: default bench line, ncu launch list, one --set full capture of render_kernel, tile timelines.
tag=${1:-r2}
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench.log 2>&1; tail -4 gpurun_out/${tag}_bench.log | cut -c1-600
if [ -f build/timeline/librrt_b200_timeline.so ] && [ -z "$NO_TIMELINE" ]; then
  export RRT_B200_LIB=$PWD/build/timeline/librrt_b200_timeline.so
  timeout 600 python tools/tile_timeline.py > gpurun_out/${tag}_timeline_4k_full.txt 2>&1
  timeout 600 python tools/tile_timeline.py --band 0 8 > gpurun_out/${tag}_timeline_4k_band0of8.txt 2>&1
  timeout 600 python tools/tile_timeline.py --camera C3 > gpurun_out/${tag}_timeline_4k_c3.txt 2>&1
  unset RRT_B200_LIB
  head -30 gpurun_out/${tag}_timeline_4k_full.txt
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/${tag}_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
ls -la gpurun_out/${tag}_prof.ncu-rep
