#!/bin/bash
# Round 2, split pipeline first contact: parity vs the fused kernel, then fused / split A/B on the bench frames.
mkdir -p gpurun_out
T=r2_15
timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{
for pipe in fused split; do
  export RRT_PIPELINE=$pipe
  run
  run --camera C3
  run --width 1920 --height 1080 --flags 3
done
} > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('media'),c.get('camera'),'ms',round(d['ms_per_step'],3),'lat',round(d.get('latency_ms_single_frame') or 0,2),'e2e ms',round(d['e2e'].get('ms_per_step',0),2), 'frac', round(d['roofline']['frac'],3), c.get('schedule','')[:100], d.get('split'))
PY
tail -5 gpurun_out/${T}_err.log
