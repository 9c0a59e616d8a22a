"""Format an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list for profiles/.
Usage: launch_list.py <launches.csv> <out.txt> "<command that was profiled>" """
import csv
import sys
from collections import defaultdict

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
hdr = rows[0]
iK, iG, iB, iV, iU = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}",
         "# (cold-cache, serialised launches: compare shares, not absolutes)", "# id, kernel, grid, block, duration_ms"]
tot = defaultdict(float)
for n, r in enumerate(rows[1:]):
    v = float(r[iV].replace(",", ""))
    ms = v / 1e6 if r[iU] in ("ns", "nsecond") else (v / 1e3 if r[iU] in ("us", "usecond") else v)
    name = r[iK][:72]
    tot[name] += ms
    lines.append(f"{n:3d}  {name:72s} {r[iG]:>16s} {r[iB]:>14s} {ms:10.3f}")
lines.append("")
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"# share {100 * v / s:6.2f} % {v:12.3f} ms  {k}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-6:]))
