#!/bin/bash
# Round 2 final: the bench line at N GPUs (split pipeline, packed tracer, peer exchange) incl. the path sub-record.
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r2_39_n${N}_err.log | tail -1 > gpurun_out/r2_39_n${N}_bench.jsonl
python - <<PY
import json
for l in open('gpurun_out/r2_39_n${N}_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['pipeline']['kind'][:5], d['config']['exchange'], 'ms', round(d['ms_per_step'],3), 'fps', round(d['frames_per_s'],1), 'e2e fps', round(d['e2e']['frames_per_s'],1), 'lat', round(d['latency_ms_single_frame'],2), d['config']['schedule'][:150], (d.get('path') or {}).get('frames_per_s'), (d.get('path') or {}).get('frames_equal_single_gpu_render'), d['gpu_launches'])
PY
tail -2 gpurun_out/r2_39_n${N}_err.log
