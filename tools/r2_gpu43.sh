#!/bin/bash
# Round 2: does the pool size (number of passes) matter on the media-heavy frames?
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 10 --warmup 6 --depth 2 --no-cpu-baseline --no-ref-cuda "$@" 2>/dev/null | tail -1; }
{
for mb in 16384 40960; do
  export RRT_POOL_MB=$mb
  run --camera C3
  timeout 600 python bench.py --workload path --steps 1 2>/dev/null | tail -1
done
} > gpurun_out/r2_43.jsonl
python - <<PY
import json
for l in open('gpurun_out/r2_43.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('camera'),'seq ms',round(d['ms_per_step'],3),'fps',round(d.get('frames_per_s',0),1),'alone',d.get('latency_ms_single_frame'), c.get('frames'), d.get('pipeline'))
PY
