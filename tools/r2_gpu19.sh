#!/bin/bash
# Round 2: split pipeline after the inline allocation fast path: parity, tile timelines (1080p; a 128x72 frame whose tiles
# run alone on their SM sub-partition), A/B.
mkdir -p gpurun_out
T=r2_19
timeout 900 python -m pytest tests/test_gpu_split.py -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
(
export RRT_B200_LIB=$PWD/build/timeline/librrt_b200_timeline.so
RRT_PIPELINE=split timeout 300 python tools/tile_timeline.py --width 1920 --height 1080 > gpurun_out/${T}_timeline_1080_split.txt 2>&1
RRT_PIPELINE=split timeout 300 python tools/tile_timeline.py --width 128 --height 72 > gpurun_out/${T}_timeline_128_split.txt 2>&1
RRT_PIPELINE=fused timeout 300 python tools/tile_timeline.py --width 128 --height 72 > gpurun_out/${T}_timeline_128_fused.txt 2>&1
RRT_PIPELINE=fused timeout 300 python tools/tile_timeline.py --width 128 --height 72 --flags 0 > gpurun_out/${T}_timeline_128_geo.txt 2>&1
)
head -24 gpurun_out/${T}_timeline_1080_split.txt
for f in 128_split 128_fused 128_geo; do echo "== $f"; sed -n 1,3p gpurun_out/${T}_timeline_$f.txt; sed -n 12,18p gpurun_out/${T}_timeline_$f.txt; done
run() { timeout 300 python bench.py --steps 10 --warmup 6 --no-cpu-baseline --no-ref-cuda "$@" 2>gpurun_out/${T}_err.log | tail -1; }
{
for pipe in fused split; do
  export RRT_PIPELINE=$pipe
  run
  run --width 1920 --height 1080 --flags 3
done
} > gpurun_out/${T}_ab.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('camera'),d['pipeline']['kind'][:5],'seq ms',round(d['ms_per_step'],3),'alone',round(d.get('latency_ms_single_frame') or 0,2),'e2e ms',round(d['e2e'].get('ms_per_step',0),2), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
PY
