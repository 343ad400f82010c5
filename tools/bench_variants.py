#!/usr/bin/env python
"""Time individual gather_reduce variants (CUDA events, median of N) on one workload.
    python tools/bench_variants.py --workload C3
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from ggad_b200 import ops, synth  # noqa: E402
from ggad_b200.graph import CSRGraph  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="S64")
ap.add_argument("--no-chase", action="store_true", help="skip the chase-mode lines (variant libraries that do not flag tiles)")
ap.add_argument("--tma-ab", action="store_true",
                help="only the plain launches, register-direct gathers vs rows staged by TMA (GGAD_TMA_ROWS / GGAD_TMA_STAGES)")
a = ap.parse_args()
n, m, d = WORKLOADS[a.workload]
g = synth.rmat_shard(n, m, seed=0)
gw = CSRGraph(g.rowptr, g.col, torch.rand(g.nnz, device="cuda"), n, n)
gw._plan = g.plan
x = torch.randn(n, d, device="cuda")
if a.tma_ab:
    print(f"workload {a.workload}: n={n} nnz={m} d={d}  (median of 10, CUDA events)")
    y_ref = ops.gather_reduce(g, x)["y"].clone()
    cfgs = [(0, 0, 3, 0, 0), (1, 2, 3, 0, 0), (2, 2, 3, 0, 0), (3, 1, 3, 0, 0), (3, 2, 3, 0, 0)]
    # kind 4: nt warps of every CTA on the TMA path, the rest register-direct; (kind, stages, promo, nt, weight/16)
    cfgs += [(4, 1, 3, 0, 16), (4, 2, 3, 1, 10), (4, 2, 3, 2, 10), (4, 2, 3, 2, 7), (4, 2, 3, 2, 13), (4, 1, 3, 2, 10), (4, 4, 3, 1, 10),
             (4, 2, 3, 3, 10), (4, 2, 3, 4, 10), (4, 3, 3, 2, 10)]
    for kind, stages, promo, nt, wq in cfgs:
        os.environ["GGAD_TMA_ROWS"], os.environ["GGAD_TMA_STAGES"] = str(kind), str(stages)
        os.environ["GGAD_TMA_L2_PROMOTION"] = str(promo)
        os.environ["GGAD_TMA_WARPS"], os.environ["GGAD_TMA_WEIGHT"] = str(nt), str(wq)
        y = ops.gather_reduce(g, x)["y"]
        same = "bit-identical" if torch.equal(y, y_ref) else f"max |diff| {float((y - y_ref).abs().max()):.2e}"
        tu = timeit(lambda: ops.gather_reduce(g, x))
        tw = timeit(lambda: ops.gather_reduce(gw, x))
        what = {0: "register-direct ld.global.nc (shipped)", 1: f"TMA tile::gather4, {stages} stages" + ("" if promo == 3 else ", no L2 promotion"),
                2: f"cp.async.bulk per row, {stages} stages",
                3: f"gather4, warp-converged, {stages} x 8-row stages",
                4: f"mixed: {nt} TMA warps/CTA x {stages} stages, weight {wq}/16"}[kind]
        print(f"  {what:48s} unweighted {tu:7.3f} ms {m / tu / 1e6:7.2f} G edges/s   weighted {tw:7.3f} ms {m / tw / 1e6:7.2f} G edges/s   {same}")
    sys.exit(0)
bias = torch.randn(d, device="cuda")
slope = torch.tensor([0.25], device="cuda")
cs = torch.rand(n, device="cuda") + 0.5
sub = torch.randperm(n, device="cuda")[: max(1, n // 7)].to(torch.int32)
res = {}
res["plain unweighted (row_scale)"] = timeit(lambda: ops.gather_reduce(g, x))
res["plain weighted"] = timeit(lambda: ops.gather_reduce(gw, x))
res["weighted + bias + PReLU + z (GCN layer)"] = timeit(lambda: ops.gather_reduce(gw, x, bias=bias, prelu_slope=slope, want_z=True))
res["weighted + bias + PReLU, no z"] = timeit(lambda: ops.gather_reduce(gw, x, bias=bias, prelu_slope=slope))
res["weighted + sumsq"] = timeit(lambda: ops.gather_reduce(gw, x, want_sumsq=True))
res["general (col_scale)"] = timeit(lambda: ops.gather_reduce(gw, x, col_scale=cs, use_graph_scales=False))
res["general, 1/7 of columns non-zero (affinity backward)"] = timeit(
    lambda: ops.gather_reduce(gw, x, col_scale=torch.zeros(n, device="cuda").index_fill_(0, sub.long(), 1.0), use_graph_scales=False))
# fused-exchange variant on one GPU: the "peer" is a local matrix, so only the kernel-side cost of the push phase shows
peer = torch.empty(n, d, device="cuda")
zero = torch.zeros(n, dtype=torch.int32, device="cuda")
some = (torch.rand(n, device="cuda") < 0.42).to(torch.int32)
res["push variant, no row needed"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=zero))
res["push variant, 42 % of rows to one local peer"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=some))
res["push variant, every row to one local peer"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()]))
seven = (torch.rand(n, device="cuda") < 0.25).to(torch.int32) * 127      # N = 8 shaped: 25 % of the rows to all 7 peers
peers7 = [peer.data_ptr()] * 7
res["push variant, 25 % of rows to 7 local peers"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=peers7, peer_need=seven))
if a.no_chase:
    print(f"workload {a.workload}: n={n} nnz={m} d={d}")
    for k, v in res.items():
        print(f"  {k:55s} {v:8.3f} ms   {m / v / 1e6:8.2f} G edges/s")
    sys.exit(0)
# chase mode: the gather only flags finished tiles, ggad_halo_chase (own stream, few CTAs) moves the rows
res["chase variant, no row needed"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=zero, chase=True))
res["chase variant, 42 % of rows to one local peer"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=some, chase=True))
res["chase variant, every row to one local peer"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], chase=True))
for ctas in (24, 96):
    res[f"chase variant, 42 % of rows, {ctas} CTAs"] = timeit(
        lambda: ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=some, chase=True, chase_ctas=ctas))
res["chase variant, 25 % of rows to 7 local peers"] = timeit(lambda: ops.gather_reduce(g, x, y_peers=peers7, peer_need=seven, chase=True))
print(f"workload {a.workload}: n={n} nnz={m} d={d}  env: " + " ".join(k for k in ("GGAD_FORCE_FULL_EPI", "GGAD_EPI_DEFERRED", "GGAD_B200_LIB") if os.environ.get(k)))
for k, v in res.items():
    print(f"  {k:55s} {v:8.3f} ms   {m / v / 1e6:8.2f} G edges/s")
