#!/usr/bin/env python
"""Tiny driver for ncu: build one workload, run forward + backward gather passes a few times.

    ncu --set full --clock-control none --import-source on -k regex:gather_tiled_kernel -s 2 -c 2 \
        -o gpurun_out/prof python tools/profile_spmm.py --workload S64 --iters 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from ggad_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="S64")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--nodes", type=int)
ap.add_argument("--edges", type=int)
ap.add_argument("--width", type=int)
ap.add_argument("--variant", default="plain", choices=["plain", "gcn_layer", "col_scale", "push", "chase", "sumsq", "hot_first"],
                help="which launch of the hot kernel to run: plain fwd+bwd; the GCN-layer epilogue (weighted + bias + PReLU "
                     "+ z); the general mode (col_scale); the fused halo push / chase exchange with a local stand-in peer")
a = ap.parse_args()
n, m, d = WORKLOADS[a.workload]
n, m, d = a.nodes or n, a.edges or m, a.width or d
g = synth.rmat_shard(n, m, seed=0)
gt = g.T
x = torch.randn(n, d, device="cuda")
if a.variant == "hot_first":
    g, gt = g.reorder_edges_hot_first(), gt.reorder_edges_hot_first()
if a.variant in ("plain", "hot_first"):          # GGAD_TMA_ROWS / GGAD_TMA_STAGES in the environment select the TMA-staged rows
    for _ in range(a.iters):
        y = ops.gather_reduce(g, x)["y"]
        dx = ops.gather_reduce(gt, y)["y"]
else:
    from ggad_b200.graph import CSRGraph
    gw = CSRGraph(g.rowptr, g.col, torch.rand(g.nnz, device="cuda"), n, n)
    gw._plan = g.plan
    bias, slope = torch.randn(d, device="cuda"), torch.tensor([0.25], device="cuda")
    cs = torch.rand(n, device="cuda") + 0.5
    peer = torch.empty(n, d, device="cuda")
    some = (torch.rand(n, device="cuda") < 0.42).to(torch.int32)
    for _ in range(a.iters):
        if a.variant == "gcn_layer":
            dx = ops.gather_reduce(gw, x, bias=bias, prelu_slope=slope, want_z=True)["y"]
        elif a.variant == "sumsq":
            dx = ops.gather_reduce(gw, x, want_sumsq=True)["y"]
        elif a.variant == "col_scale":
            dx = ops.gather_reduce(gw, x, col_scale=cs, use_graph_scales=False)["y"]
        elif a.variant == "push":
            dx = ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=some)["y"]
        else:
            dx = ops.gather_reduce(g, x, y_peers=[peer.data_ptr()], peer_need=some, chase=True)["y"]
torch.cuda.synchronize()
print("done", float(dx[0, 0]))
