"""Static view of a kernel's loops from cuobjdump SASS: instruction count and opcode mix per loop.
Usage: sass_loops.py <lib.so> <function-substring>"""
import re
import subprocess
import sys
from collections import Counter

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on, L = False, []
for line in txt.splitlines():
    if "Function :" in line:
        on = key in line
        if on:
            print(line.strip())
    elif on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            L.append((int(m.group(1), 16), m.group(2).strip()))
print("total instructions:", len(L))


def opc(s):
    t = s.split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]


for a, s in L:
    m = re.search(r"\bBRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", s)
    if m and int(m.group(1), 16) < a:
        t = int(m.group(1), 16)
        body = [x for x in L if t <= x[0] <= a]
        c = Counter(opc(x[1]) for x in body)
        print(f"loop {t:#x}..{a:#x}: {len(body)} instrs", dict(c.most_common(16)))
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    for a, s in L:
        if lo <= a <= hi:
            print(f"{a:#06x}  {s}")
