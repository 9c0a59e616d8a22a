#!/bin/bash
# Round 2: what slows the centre tiles of trace_kernel?  One knob per build, 1080p frame alone (RRT_PIPELINE=split).
mkdir -p gpurun_out
for v in default vb vc vd ve vf vg; do
  if [ $v = default ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/$v/librrt_$v.so; fi
  echo "== $v"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --width 1920 --height 1080 --reps 4 2>&1 | tail -2 | cut -c1-40
done
