#!/bin/bash
# Round 2 final profile session (split pipeline, packed tracer): default bench line, ncu launch list, --set full captures of
# trace_kernel_p and media_kernel, the other BASELINE configs, tile timelines of the trace kernel.
mkdir -p gpurun_out
T=r2_38
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.log 2>&1; tail -4 gpurun_out/${T}_bench.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --depth 2 --no-cpu-baseline --no-ref-cuda > gpurun_out/${T}_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel_p --launch-skip 2 --launch-count 1 \
  -f -o gpurun_out/${T}_trace python tools/render_once.py --reps 3 > gpurun_out/${T}_ncu_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:media_kernel --launch-skip 2 --launch-count 1 \
  -f -o gpurun_out/${T}_media python tools/render_once.py --reps 3 > gpurun_out/${T}_ncu_media.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:fold_kernel --launch-skip 2 --launch-count 1 \
  -f -o gpurun_out/${T}_fold python tools/render_once.py --reps 3 > gpurun_out/${T}_ncu_fold.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda "$@" 2>/dev/null | tail -1; }
{
run --width 256 --height 256 --flags 0
run --width 1920 --height 1080 --flags 1
run --width 1920 --height 1080 --flags 3
run --camera C3
run --camera C1
RRT_PIPELINE=fused run
timeout 600 python bench.py --workload path --steps 1 2>/dev/null | tail -1
} > gpurun_out/${T}_configs.jsonl
python - <<PY
import json
for l in open('gpurun_out/${T}_configs.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']
        print(c.get('width'),c.get('height'),c.get('media'),c.get('camera'),(d.get('pipeline') or {}).get('kind','')[:5],'ms',round(d['ms_per_step'],3),'fps',round(d.get('frames_per_s',0),1),'steps/s %.3e'%d['value'],'lat',d.get('latency_ms_single_frame'), c.get('frames'))
PY
export RRT_B200_LIB=$PWD/build/timeline/librrt_b200_timeline.so
RRT_PIPELINE=split RRT_TRACE=scalar timeout 300 python tools/tile_timeline.py > gpurun_out/${T}_timeline_4k_split_trace.txt 2>&1
RRT_PIPELINE=split RRT_TRACE=scalar timeout 300 python tools/tile_timeline.py --band 0 8 > gpurun_out/${T}_timeline_4k_split_trace_band0of8.txt 2>&1
head -12 gpurun_out/${T}_timeline_4k_split_trace.txt
head -12 gpurun_out/${T}_timeline_4k_split_trace_band0of8.txt
