#!/bin/bash
# Round 2: per-kernel durations / instruction counts / lane efficiency of the split pipeline (ncu, serialised launches).
mkdir -p gpurun_out
T=r2_16
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for cam in C0 C3; do
  for pipe in split fused; do
    RRT_PIPELINE=$pipe timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|media_kernel|fold_kernel|sweep_kernel|render_kernel' -c 80 --csv \
      --log-file gpurun_out/${T}_${cam}_${pipe}.csv python bench.py --steps 1 --warmup 1 --depth 1 --camera $cam --no-cpu-baseline --no-ref-cuda > gpurun_out/${T}_${cam}_${pipe}.log 2>&1
    echo "$cam $pipe rc=$?"
  done
done
python - <<'PY'
import csv, glob, collections
for f in sorted(glob.glob('gpurun_out/r2_16_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if not rows: print(f,'empty'); continue
    hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
    per=collections.OrderedDict()
    for r in rows[1:]:
        key=(r[idx['ID']], r[idx['Kernel Name']][:40])
        per.setdefault(key,{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
    print('==',f)
    for (i,k),m in list(per.items())[-40:]:
        print(i,k,' '.join(f"{n.split('.')[0][-28:]}={v}" for n,v in m.items()))
PY
