#!/bin/bash
# Round 2: per-kernel view of the split pipeline at 1080p, and of the geodesic-only kernels at 4K (emission overhead).
mkdir -p gpurun_out
T=r2_17
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg
K='trace_kernel|media_kernel|fold_kernel|sweep_kernel|render_kernel'
run() { tag=$1; shift; timeout 600 ncu --metrics $M --clock-control none -k regex:"$K" -c 40 --csv --log-file gpurun_out/${T}_$tag.csv python bench.py --steps 1 --warmup 1 --depth 1 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${T}_$tag.log 2>&1; echo "$tag rc=$?"; }
RRT_PIPELINE=split run 1080_split --width 1920 --height 1080
RRT_PIPELINE=fused run 1080_fused --width 1920 --height 1080
RRT_KERNEL=scalar run geo_scalar --flags 0
RRT_KERNEL=packed run geo_packed --flags 0
python - <<'PY'
import csv, glob, collections
for f in sorted(glob.glob('gpurun_out/r2_17_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if not rows: print(f,'empty'); continue
    hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
    per=collections.OrderedDict()
    for r in rows[1:]:
        key=(r[idx['ID']], r[idx['Kernel Name']][:34])
        per.setdefault(key,{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
    print('==',f)
    for (i,k),m in list(per.items())[-9:]:
        print(i,k,' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
# sequences without ncu: depth 2 and depth 1, both pipelines, 1080p
for pipe in fused split; do for d in 1 2; do
  RRT_PIPELINE=$pipe timeout 300 python bench.py --steps 20 --warmup 5 --depth $d --width 1920 --height 1080 --no-cpu-baseline --no-ref-cuda 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$pipe depth $d', 'seq ms=%.2f'%d['ms_per_step'], 'alone=%.2f'%d['latency_ms_single_frame'], 'e2e=%.2f'%d['e2e']['ms_per_step'], d.get('pipeline'))"
done; done
