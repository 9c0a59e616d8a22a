#!/bin/bash
# bench at N GPUs (as the driver launches it).  Usage: tools/gpu_scale.sh <tag> <N> [bench args]
tag=$1; n=$2; shift 2
if [ "$n" = 1 ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/${tag}_n1.log 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 "$@" > gpurun_out/${tag}_n$n.log 2>&1
fi
tail -1 gpurun_out/${tag}_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'ms=%.2f'%d['ms_per_step'], 'fps=%.1f'%d['frames_per_s'], 'lat=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'kernel_ms=%.2f'%d['roofline']['kernel_ms'], 'frac=%.3f'%d['roofline']['frac'], 'e2e_fps=%.1f'%d['e2e']['frames_per_s'], d['clocks']['sm_mhz'])" || tail -20 gpurun_out/${tag}_n$n.log
