#!/usr/bin/env python
"""Summarise an ncu report (raw page) into the metrics DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass-top N]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    name_i = hdr.index("Kernel Name")
    print("| metric | unit | " + " | ".join(r[name_i].split("(")[0][-40:] for r in rows) + " |")
    print("|---|---|" + "---|" * len(rows))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows) + " |")
    if "--sass-top" in sys.argv:
        n = int(sys.argv[sys.argv.index("--sass-top") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        secs, cur = [], None
        for r in csv.reader(io.StringIO(out)):
            if r and r[0] == "Kernel Name":
                cur = dict(name=r[1], rows=[])
                secs.append(cur)
            elif r and r[0] == "Address":
                cur["hdr"] = r
            elif cur is not None and r:
                cur["rows"].append(r)
        for s in secs[:1]:
            h = s["hdr"]
            ie, is_ = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
            tot = sum(int(r[ie]) for r in s["rows"])
            tots = sum(int(r[is_]) for r in s["rows"])
            print(f"\nSASS of {s['name']}: {len(s['rows'])} instructions, {tot} warp-instructions executed, {tots} stall samples")
            c, cs = Counter(), Counter()
            for r in s["rows"]:
                t = r[1].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                c[op] += int(r[ie])
                cs[op] += int(r[is_])
            for op, v in c.most_common(14):
                print(f"  {op:8s} {v / tot * 100:5.1f}% of executed  {cs[op] / max(tots, 1) * 100:5.1f}% of stall samples")
            print("  hottest instructions by stall samples:")
            for r in sorted(s["rows"], key=lambda r: -int(r[is_]))[:n]:
                print(f"    {int(r[is_]) / max(tots, 1) * 100:5.1f}%  exec={r[ie]:>10s}  {r[1][:80]}")


if __name__ == "__main__":
    main()
