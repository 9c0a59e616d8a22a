#!/bin/bash
# Round 2, 8-GPU session: multirank checks at N=8 (both exchange paths, path sequence), bench at N=8 (peer, nccl) and N=4.
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_11_ngpus.txt
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
{ tr 8 29701 tests/tools/check_bands_multirank.py; tr 8 29702 tests/tools/check_path_multirank.py; } > gpurun_out/r2_11_checks.log 2>&1
grep -E "multirank ok|PathSequence|Error|assert" gpurun_out/r2_11_checks.log | head
{
echo "== n8 peer"; tr 8 29703 bench.py --gpus 8 --steps 20 --warmup 5 --exchange peer 2>gpurun_out/r2_11_err1.log | tail -1
echo "== n8 nccl"; tr 8 29704 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-path 2>gpurun_out/r2_11_err2.log | tail -1
echo "== n4 peer"; tr 4 29705 bench.py --gpus 4 --steps 20 --warmup 5 --exchange peer 2>gpurun_out/r2_11_err3.log | tail -1
} > gpurun_out/r2_11_bench.jsonl
python - <<PY
import json
for l in open('gpurun_out/r2_11_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['n_gpus'], d['config']['exchange'], 'ms', round(d['ms_per_step'],2), 'fps', round(d['frames_per_s'],1), 'e2e fps', round(d['e2e']['frames_per_s'],1), 'lat', round(d['latency_ms_single_frame'],2), 'kern', round(d['roofline']['kernel_ms'],2), d['config']['schedule'][-150:], (d.get('path') or {}).get('frames_per_s'), (d.get('path') or {}).get('frames_equal_single_gpu_render'))
    else: print(l.strip())
PY
tail -3 gpurun_out/r2_11_err1.log
