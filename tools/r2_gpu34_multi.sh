#!/bin/bash
# Round 2, multi-GPU session with the split pipeline: multirank checks (N = 2) or the bench line (N given), peer exchange.
mkdir -p gpurun_out
N=${1:-2}
T=r2_34_n$N
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/${T}_pytest_multirank.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_multirank.log
  tail -4 gpurun_out/${T}_pytest_multirank.log
fi
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>gpurun_out/${T}_bench_err.log | tail -1
}
{
echo "== split peer"; run --exchange peer
echo "== fused peer"; RRT_PIPELINE=fused run --exchange peer --no-path
if [ "$N" != "8" ]; then echo "== split peer depth 2"; run --exchange peer --no-path --depth 2; fi
} > gpurun_out/${T}_bench.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['pipeline']['kind'][:5], d['config']['exchange'], 'ms', round(d['ms_per_step'],2), 'fps', round(d['frames_per_s'],1), 'e2e fps', round(d['e2e']['frames_per_s'],1), 'lat', round(d['latency_ms_single_frame'],2), d['config']['schedule'][:150], (d.get('path') or {}).get('frames_per_s'), (d.get('path') or {}).get('frames_equal_single_gpu_render'))
    else: print(l.strip())
PY
tail -3 gpurun_out/${T}_bench_err.log
