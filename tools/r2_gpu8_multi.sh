#!/bin/bash
# Round 2, multi-GPU session (N = 2): band / path multirank checks through both exchange paths, bench at N=2 (both exchanges).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_8_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r2_8_pytest_multirank.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_8_pytest_multirank.log
tail -15 gpurun_out/r2_8_pytest_multirank.log
N=${1:-2}
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>gpurun_out/r2_8_bench_err.log | tail -1
}
{
echo "== peer"; run --exchange peer
echo "== nccl"; run --exchange nccl --no-path
echo "== reference arm"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1
} > gpurun_out/r2_8_bench_n$N.jsonl
python - <<PY
import json
for l in open('gpurun_out/r2_8_bench_n$N.jsonl'):
    if l.startswith('{'):
        d=json.loads(l)
        if d.get('impl')=='reference': print('ref', d['value'], d['cpu_baseline']['cores'], d['config']['wall_s']); continue
        print(d['config']['exchange'], 'ms', round(d['ms_per_step'],2), 'fps', round(d['frames_per_s'],1), 'e2e fps', round(d['e2e']['frames_per_s'],1), 'lat', round(d['latency_ms_single_frame'],2), d['config']['schedule'][:120], d.get('path'))
    else: print(l.strip())
PY
tail -5 gpurun_out/r2_8_bench_err.log
