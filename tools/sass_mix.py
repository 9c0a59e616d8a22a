"""Aggregate an `ncu --page source --csv` export: executed warp-instructions per opcode and per code
region (split at a given address list), to see where issue slots go.  Usage: sass_mix.py file.csv [steps]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
iA, iS, iE, iT = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ops, tot, thr = Counter(), 0, 0
lines = []
for r in rows[2:]:
    if len(r) <= iE or not r[iE]:
        continue
    n = int(r[iE]); t = int(r[iT])
    src = r[iS].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += n
    tot += n; thr += t
    lines.append((r[iA], n, src))
print(f"total warp-inst {tot:,}  thread-inst/warp-inst {thr/tot:.2f}")
norm = (steps / 32.0) if steps else None
for op, n in ops.most_common(40):
    print(f"{op:12s} {n:>16,} {100*n/tot:6.2f}%" + (f"  {n/norm:8.2f}/warp-step" if norm else ""))
if norm:
    print(f"total per warp-step: {tot/norm:.1f}")
if len(sys.argv) > 3:
    # dump hot lines
    for a, n, s in sorted(lines, key=lambda x: -x[1])[: int(sys.argv[3])]:
        print(a, f"{n:,}", s)
