#!/usr/bin/env python
"""Time the dense projections of the GGAD path: ggad_dense_matmul (tcgen05 fp32-accurate GEMM / SIMT) next to
torch's fp32 matmul (cuBLAS, TF32 off) on the shapes programs A and B run.
    python tools/bench_dense.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggad_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, iters=20, warm=5):
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


shapes = [("C1 layer 1: X W^T", 7535, 300, 748), ("C1 layer 2", 7535, 300, 300), ("C3 layer 2", 39357, 300, 300),
          ("C3 layer 1 (d=10->12)", 39357, 300, 12), ("MLP fc1", 1300, 152, 300), ("mini-batch |U| x h x d", 3000, 64, 20),
          ("C4 full-graph h=64", 3_700_550, 64, 20)]
print(f"{'shape':32s} {'M':>8s} {'N':>5s} {'K':>5s} {'tcgen05 ms':>11s} {'TFLOP/s':>8s} {'SIMT ms':>9s} {'cuBLAS ms':>10s}  max rel err (tc / simt / cublas)")
for name, m, n, k in shapes:
    x, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    ref = (x[:4096].double() @ w.double().t())
    res = {}
    for tag, fn in (("tc", lambda: ops.dense_matmul(x, w, trans_b=True, path=2)), ("simt", lambda: ops.dense_matmul(x, w, trans_b=True, path=1)),
                    ("cublas", lambda: x @ w.t())):
        t = timeit(fn)
        err = ((fn()[:4096].double() - ref).abs().max() / ref.abs().max()).item()
        res[tag] = (t, err)
    fl = 2.0 * m * n * k
    print(f"{name:32s} {m:8d} {n:5d} {k:5d} {res['tc'][0]:11.4f} {fl / res['tc'][0] / 1e9:8.2f} {res['simt'][0]:9.4f} {res['cublas'][0]:10.4f}"
          f"  {res['tc'][1]:.2e} / {res['simt'][1]:.2e} / {res['cublas'][1]:.2e}")
