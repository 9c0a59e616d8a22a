"""Turn an .ncu-rep of render_kernel into the small text summary kept under profiles/.
Usage: ncu_summary.py <rep> <rk4_steps_in_launch> <out.txt>"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, steps, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
lines = [f"ncu summary of {rep}", ""]
for k in keys:
    if k in m:
        lines.append(f"{k:75s} {m[k][0]:>22s} {m[k][1]}")
lines.append("")
lines.append("warp stall reasons (avg warps per issue-active cycle):")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = float(m[h][0].replace(",", ""))
        if v >= 0.01:
            lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
iS, iE, iT = sh.index("Source"), sh.index("Instructions Executed"), sh.index("Thread Instructions Executed")
ops, tot, thr = Counter(), 0, 0
for r in srows[2:]:
    if len(r) <= iE or not r[iE]:
        continue
    n = int(r[iE])
    s = r[iS].strip().split()
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    ops[op] += n
    tot += n
    thr += int(r[iT])
lines += ["", f"executed warp instructions: {tot:,}  ({tot / (steps / 32):.1f} per warp-step over {steps:.0f} RK4 steps; "
          f"{thr / tot:.2f} active threads per instruction)", "opcode mix (warp instructions per warp-step):"]
for op, n in ops.most_common(18):
    lines.append(f"  {op:10s} {n / (steps / 32):8.2f}  {100 * n / tot:5.1f} %")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
