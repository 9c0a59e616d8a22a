#!/bin/bash
# Round 2, split pipeline v2 at 1080p: per-kernel times and the trace kernel's tile timeline.
mkdir -p gpurun_out
T=r2_23
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active
RRT_PIPELINE=split timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|media_kernel|fold_kernel|sweep_kernel|render_kernel' -c 12 --csv \
   --log-file gpurun_out/${T}_1080_split.csv python tools/render_once.py --width 1920 --height 1080 --reps 3 > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_23_1080_split.csv')) if len(r)>10]
hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idx['ID']], r[idx['Kernel Name']][:30]),{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
for (i,k),m in list(per.items())[-4:]:
    print(i,k,' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
export RRT_B200_LIB=$PWD/build/timeline/librrt_b200_timeline.so
RRT_PIPELINE=split timeout 300 python tools/tile_timeline.py --width 1920 --height 1080 > gpurun_out/${T}_timeline_1080_split.txt 2>&1
head -32 gpurun_out/${T}_timeline_1080_split.txt
