#!/usr/bin/env python
"""Time the dense projections of the GGAD path: ggad_dense_matmul (tcgen05 fp32-accurate GEMM / SIMT, and what the
auto dispatch picks) next to torch's fp32 matmul (cuBLAS, TF32 off) on the shapes programs A and B run -- forward
(x W^T), data gradient (dy W) and weight gradient (dy^T x).
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


shapes = [("C1 layer 1", 7535, 300, 748), ("C1 layer 2", 7535, 300, 300), ("C2 layer 1 (d=25->28)", 11944, 300, 28),
          ("C2 layer 2", 11944, 300, 300), ("C3 layer 1 (d=10->12)", 39357, 300, 12), ("C3 layer 2", 39357, 300, 300),
          ("MLP fc1", 1300, 152, 300), ("MLP fc2", 1300, 76, 152), ("mini-batch |U| x h x d", 8000, 64, 20),
          ("C4 full-graph h=64", 3_700_550, 64, 20)]
print(f"{'shape (M nodes, N out, K in)':30s} {'op':7s} {'auto ms':>8s} {'tcgen05':>8s} {'SIMT':>8s} {'cuBLAS':>8s}   rel err auto / cublas")
for name, m, n, k in shapes:
    x, w, dy = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda"), torch.randn(m, n, device="cuda")
    ops_ = [("x W^T", lambda p: ops.dense_matmul(x, w, trans_b=True, path=p), lambda: x @ w.t(), lambda: x[:4096].double() @ w.double().t(), slice(0, 4096)),
            ("dy W", lambda p: ops.dense_matmul(dy, w, path=p), lambda: dy @ w, lambda: dy[:4096].double() @ w.double(), slice(0, 4096)),
            ("dy^T x", lambda p: ops.dense_matmul(dy, x, trans_a=True, path=p), lambda: dy.t() @ x, lambda: dy.double().t() @ x.double(), slice(None))]
    for tag, ours, cublas, ref_fn, sl in ops_:
        ref = ref_fn()
        t = {}
        for key, p in (("auto", 0), ("tc", 2), ("simt", 1)):
            try:
                t[key] = timeit(lambda: ours(p))
            except RuntimeError:
                t[key] = float("nan")
        t["cublas"] = timeit(cublas)
        e_auto = ((ours(0)[sl].double() - ref).abs().max() / ref.abs().max()).item()
        e_cub = ((cublas()[sl].double() - ref).abs().max() / ref.abs().max()).item()
        print(f"{name:30s} {tag:7s} {t['auto']:8.4f} {t['tc']:8.4f} {t['simt']:8.4f} {t['cublas']:8.4f}   {e_auto:.1e} / {e_cub:.1e}")
