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
a = ap.parse_args()
n, m, d = WORKLOADS[a.workload]
n, m, d = a.nodes or n, a.edges or m, a.width or d
g = synth.rmat_shard(n, m, seed=0)
gt = g.T
x = torch.randn(n, d, device="cuda")
for _ in range(a.iters):
    y = ops.gather_reduce(g, x)["y"]
    dx = ops.gather_reduce(gt, y)["y"]
torch.cuda.synchronize()
print("done", float(dx[0, 0]))
