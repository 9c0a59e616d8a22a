#!/bin/bash
mkdir -p gpurun_out
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in k0 default; do ab $lib; done
for lib in k0 default; do ab $lib --camera C3; done
for lib in k0 default; do ab $lib --camera C1; done
for lib in k0 default; do ab $lib --camera C2; done
for lib in k0 default; do ab $lib --flags 0; done
} 2>&1 | tee gpurun_out/r2_6_ab.log
unset RRT_B200_LIB
# where do the C3 cycles go: executed-instruction census of one C3 launch, both builds
for lib in k0 default; do
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fma.sum,smsp__cycles_active.sum,smsp__issue_active.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:render_kernel --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r2_6_ncu_c3_$lib.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 --camera C3 > gpurun_out/r2_6_ncu_c3_$lib.log 2>&1
done
