#!/bin/bash
# Round 2: the other BASELINE configs on one B200 with the default bench options, smoke, full GPU suite.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_14_smoke.log 2>&1; tail -1 gpurun_out/r2_14_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_14_pytest.log
tail -4 gpurun_out/r2_14_pytest.log
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda "$@" 2>/dev/null | tail -1; }
{
run --width 256 --height 256 --flags 0
run --width 1920 --height 1080 --flags 1
run --width 1920 --height 1080 --flags 3
run --camera C3
run --camera C1
timeout 600 python bench.py --workload path --steps 1 2>/dev/null | tail -1
} > gpurun_out/r2_14_configs.jsonl
python - <<PY
import json
for l in open('gpurun_out/r2_14_configs.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('media'),c.get('camera'),'ms',round(d['ms_per_step'],3),'fps',round(d.get('frames_per_s',0),1),'steps/s %.3e'%d['value'],'lat',d.get('latency_ms_single_frame'), c.get('schedule','')[:60], c.get('frames'))
PY
