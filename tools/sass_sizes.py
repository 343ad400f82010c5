#!/usr/bin/env python
"""Per-kernel SASS size / register / stack table of an object file or the shared library (cuobjdump).
    python tools/sass_sizes.py ggad_b200/build/gather_inst_g16c1.o [filter]
"""
import re
import subprocess
import sys

path = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
usage = {}
name = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
    if m and name:
        usage[name] = tuple(int(t) for t in m.groups())
        name = None
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
counts, cur = {}, None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = 0
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        counts[cur] += 1
for k in sorted(counts):
    if flt and flt not in k:
        continue
    dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(ggad::GatherArgs\)|ggad::|void ", "", dem)
    r = usage.get(k, (0, 0, 0))
    print(f"{counts[k]:6d} instr {counts[k] * 16 / 1024:6.1f} KB  reg {r[0]:3d} stack {r[1]:3d}  {dem}")
