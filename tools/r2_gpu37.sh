#!/bin/bash
# Round 2: media kernel with the batch fetched into shared memory up front: parity, per-kernel times, bench lines.
mkdir -p gpurun_out
T=r2_37
timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|media_kernel|fold_kernel|sweep_kernel|render_kernel' -c 8 --csv \
   --log-file gpurun_out/${T}_c0.csv python tools/render_once.py --reps 2 > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_37_c0.csv')) if len(r)>10]
hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idx['ID']], r[idx['Kernel Name']][:30]),{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
for (i,k),m in list(per.items())[-4:]:
    print(i,k,' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{ run; run --camera C3; run --width 1920 --height 1080 --flags 3; run --width 1920 --height 1080 --flags 1; timeout 600 python bench.py --workload path --steps 1 2>/dev/null | tail -1; } > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('media'),c.get('camera'),'seq ms',round(d['ms_per_step'],3),'fps',round(d.get('frames_per_s',0),1),'alone',d.get('latency_ms_single_frame'),'frac', (d.get('roofline') or {}).get('frac'), c.get('frames'))
PY
