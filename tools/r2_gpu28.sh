#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for v in default dbg4; do
if [ $v = default ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/$v/librrt_$v.so; fi
RRT_PIPELINE=split timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|media_kernel|fold_kernel|sweep_kernel' -c 12 --csv \
   --log-file gpurun_out/r2_28_128_$v.csv python tools/render_once.py --width 128 --height 72 --reps 3 > gpurun_out/r2_28_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_28_128_$v.csv')) if len(r)>10]
hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idx['ID']], r[idx['Kernel Name']][:30]),{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
print("== $v")
for (i,k),m in list(per.items())[-4:]:
    print(i,k,' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
done
