#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed_op_branch.sum,sm__cycles_active.avg,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_sleeping_per_warp_active.pct,smsp__warp_issue_stalled_misc_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_drain_per_warp_active.pct,smsp__warp_issue_stalled_imc_miss_per_warp_active.pct
run() { tag=$1; shift; timeout 600 ncu --metrics $M --clock-control none -k regex:'trace_kernel|render_kernel' -c 3 --csv --log-file gpurun_out/r2_30_$tag.csv "$@" > gpurun_out/r2_30_ncu.log 2>&1; }
RRT_PIPELINE=split run split python tools/render_once.py --width 128 --height 72 --reps 3
RRT_KERNEL=scalar run geo python tools/render_once.py --width 128 --height 72 --flags 0 --reps 3
RRT_B200_LIB=$PWD/build/dbg3/librrt_dbg3.so RRT_PIPELINE=split run dbg3 python tools/render_once.py --width 128 --height 72 --reps 3
python - <<'PY'
import csv, collections
for tag in ('split','geo','dbg3'):
    rows=[r for r in csv.reader(open(f'gpurun_out/r2_30_{tag}.csv')) if len(r)>10]
    hdr=rows[0]; idx={n:i for i,n in enumerate(hdr)}
    per=collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[idx['ID']], r[idx['Kernel Name']][:26]),{})[r[idx['Metric Name']]]=r[idx['Metric Value']]
    (i,k),m=list(per.items())[-1]
    print('==',tag,k)
    for n,v in m.items():
        try: fv=float(v.replace(',',''))
        except: fv=0
        if fv!=0: print('   ',n.replace('smsp__warp_issue_stalled_','stall ').replace('_per_warp_active.pct',''),v)
PY
