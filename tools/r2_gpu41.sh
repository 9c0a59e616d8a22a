#!/bin/bash
# Round 2: zone bursts in the packed tracer + gated disk density in media_kernel: parity (both tracers), bench lines.
mkdir -p gpurun_out
T=r2_41
timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{ run; run --camera C3; run --camera C1; run --width 1920 --height 1080 --flags 3; timeout 600 python bench.py --workload path --steps 1 2>/dev/null | tail -1; } > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('media'),c.get('camera'),'seq ms',round(d['ms_per_step'],3),'fps',round(d.get('frames_per_s',0),1),'alone',d.get('latency_ms_single_frame'),'frac', (d.get('roofline') or {}).get('frac'), c.get('frames'), (d.get('pipeline') or {}).get('passes_per_frame'))
PY
