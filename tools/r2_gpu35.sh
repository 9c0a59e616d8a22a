#!/bin/bash
# Round 2: packed tracer in the split pipeline: parity (both tracers), A/B.
mkdir -p gpurun_out
T=r2_35
for tr in packed scalar; do
RRT_TRACE=$tr timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest_$tr.log 2>&1; echo "pytest $tr rc=$?" >> gpurun_out/${T}_pytest_$tr.log
tail -3 gpurun_out/${T}_pytest_$tr.log
done
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{
for tr in scalar packed; do
  export RRT_TRACE=$tr
  run
  run --width 1920 --height 1080 --flags 3
  run --camera C3
  run --camera C1
done
} > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
i=0
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(['scalar','packed'][i//4], c.get('width'),c.get('height'),c.get('camera'),'seq ms',round(d['ms_per_step'],3),'alone',round(d.get('latency_ms_single_frame') or 0,2),'e2e ms',round(d['e2e'].get('ms_per_step',0),2), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['pipeline'].get('passes_per_frame'))
        i+=1
PY
