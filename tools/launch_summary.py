#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
    python tools/launch_summary.py profiles/r01_launches_bench_S64.csv [skip_first_n]
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = defaultdict(lambda: [0, 0.0])
for r in rows[skip:]:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"<.*", "<...>", name) if "cub::" in name or "at::" in name else name
    t = float(r[14].replace(",", ""))
    unit = r[13]
    t_us = t / 1000.0 if unit in ("ns", "nsecond") else (t * 1000.0 if unit in ("ms", "msecond") else t)
    agg[name][0] += 1
    agg[name][1] += t_us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows) - skip} launches, {tot / 1000:.3f} ms of kernel time (serialised, cold-cache: compare shares)\n")
print("| kernel | launches | total us | share |")
print("|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {t:.1f} | {t / tot * 100:.1f}% |")
