#!/bin/bash
# Round 2, split pipeline v2 (row-structured sample streams, no per-step allocation): parity, per-kernel times, A/B.
mkdir -p gpurun_out
T=r2_32
timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
echo "== split 1080p alone"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --width 1920 --height 1080 --reps 4 2>&1 | tail -2 | cut -c1-200
echo "== split 4K alone"; RRT_PIPELINE=split timeout 200 python tools/render_once.py --reps 4 2>&1 | tail -2 | cut -c1-200
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
RRT_PIPELINE=split timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|media_kernel|fold_kernel|sweep_kernel|render_kernel' -c 12 --csv \
   --log-file gpurun_out/${T}_c0_split.csv python tools/render_once.py --reps 3 > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_32_c0_split.csv')) if len(r)>10]
hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idx['ID']], r[idx['Kernel Name']][:30]),{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
for (i,k),m in list(per.items())[-8:]:
    print(i,k,' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{
for pipe in fused split; do
  export RRT_PIPELINE=$pipe
  run
  run --width 1920 --height 1080 --flags 3
  run --camera C3
done
} > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('camera'),d['pipeline']['kind'][:5],'seq ms',round(d['ms_per_step'],3),'alone',round(d.get('latency_ms_single_frame') or 0,2),'e2e ms',round(d['e2e'].get('ms_per_step',0),2), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['pipeline'].get('passes_per_frame'))
PY
