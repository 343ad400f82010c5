#!/usr/bin/env python
"""One full-batch GGAD training epoch (program A: run.py:145-215) on graphs of the BASELINE.json shapes.

    python tools/bench_epoch.py --config C1|C2|C3 [--epochs 20] [--cpu-epochs 1]

GPU path: ggad_b200.model.Model + ggad_b200.losses.ggad_loss + backward + Adam (CUDA events, median).
CPU path (same box, all host threads): the oracle's CSR restatement of the same epoch (O(nnz)); the
reference's own dense N x N epoch is measured by the survey on Photo shape only (5.46 s / epoch on 8 vCPU,
BASELINE.md) -- it needs >= 4 dense N^2 fp32 tensors and does not fit at T-Finance size.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {
    # name: (nodes, undirected edges, feature width, anomaly rate, outlier fraction dataset key, noise mean/var)
    "C1": (7_535, 119_043, 745, 0.092, "photo", (0.02, 0.01)),
    "C2": (11_944, 2_199_196, 25, 0.069, "Amazon", (0.0, 0.0)),
    "C3": (39_357, 10_611_271, 10, 0.046, "t_finance", (0.0, 0.0)),
}


def synth_graph(n, m_undirected, d, rate, seed=0):
    """Symmetric binary power-law graph with ~2*m stored entries + planted anomalies."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    w = (1.0 - rng.random(n)) ** (-1.0 / 1.5)
    p = w / w.sum()
    src = rng.choice(n, int(m_undirected * 1.08), p=p)
    dst = rng.integers(0, n, len(src))
    keep = src != dst
    a = sp.coo_matrix((np.ones(keep.sum(), np.float32), (src[keep], dst[keep])), shape=(n, n)).tocsr()
    a = ((a + a.T) > 0).astype(np.float64).tocsr()
    labels = (rng.random(n) < rate).astype(np.int64)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[labels == 1] += 1.5 * rng.standard_normal((int(labels.sum()), d)).astype(np.float32)
    return a, x, labels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C1", choices=list(CONFIGS))
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--cpu-epochs", type=int, default=1)
    ap.add_argument("--h", type=int, default=300)
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay part (for kernel launch lists)")
    args = ap.parse_args()
    import random

    import scipy.sparse as sp

    import oracle                                           # only for the timed CPU leg below
    from ggad_b200 import _lib, data, graph, losses, model
    n, m, d, rate, ds, (mean, var) = CONFIGS[args.config]
    a, x, labels = synth_graph(n, m, d, rate)
    if ds in ("Amazon",):                                   # run.py:87-88 row-normalises these
        x = np.asarray(data.preprocess_features(sp.csr_matrix(np.abs(x)))[0], dtype=np.float32)
    _, _, _, idx_test, normal, abnormal = data.semi_supervised_split(labels, ds, rng=random.Random(0))
    h = args.h
    torch.manual_seed(0)
    m_gpu = model.Model(d, h, "prelu", 1, "avg")
    state = {k: v.detach().clone() for k, v in m_gpu.state_dict().items()}
    m_gpu = m_gpu.cuda()
    g_hat, g_r = graph.full_batch_graphs(a, "cuda")
    opt = torch.optim.Adam(m_gpu.parameters(), lr=1e-3)
    xt = torch.from_numpy(x).unsqueeze(0).cuda()
    ns = types.SimpleNamespace(mean=mean, var=var)
    gen = torch.Generator().manual_seed(1)

    def epoch():
        noise = (torch.randn(1, len(abnormal), h, generator=gen) * var + mean).cuda()
        opt.zero_grad(set_to_none=True)
        emb, comb, logits, emb_con, emb_abn = m_gpu(xt, g_hat, abnormal, normal, True, ns, noise=noise)
        loss = losses.ggad_loss(emb, logits, emb_con, emb_abn, g_r, normal, abnormal)[0]
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        epoch()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    ts = []
    for _ in range(args.epochs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        loss = epoch()
        e1.record()
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    launches = (_lib.launch_count() - l0) / args.epochs
    gpu_ms = float(np.median([t[0] for t in ts]))
    wall_ms = float(np.median([t[1] for t in ts]))

    # same epoch replayed as one CUDA graph (ggad_b200.train.GraphedFullBatchStep)
    from ggad_b200 import train
    m_g = model.Model(d, h, "prelu", 1, "avg")
    m_g.load_state_dict(state)
    stepper = None if args.no_graph else train.GraphedFullBatchStep(m_g.cuda(), xt, g_hat, g_r, normal, abnormal, ns)
    tg = [(float("nan"), float("nan"))] * 4 if args.no_graph else []
    for _ in range(0 if args.no_graph else args.epochs + 3):
        noise = torch.randn(1, len(abnormal), h, generator=gen) * var + mean
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        stepper.step(noise)
        e1.record()
        torch.cuda.synchronize()
        tg.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    graph_ms = float(np.median([t[0] for t in tg[3:]]))
    graph_wall_ms = float(np.median([t[1] for t in tg[3:]]))

    cpu_s = None
    if args.cpu_epochs > 0 and g_hat.nnz * h * 4 < 2e9:   # the oracle materialises [nnz, h] messages
        p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state.items()}
        a_hat_cpu, r_cpu = oracle.build_full_batch_graph(a)
        a_hat_cpu, r_cpu = oracle.csr_arrays(a_hat_cpu), oracle.csr_arrays(r_cpu)
        opt_c = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-3)
        tt = []
        for _ in range(args.cpu_epochs):
            t0 = time.perf_counter()
            opt_c.zero_grad()
            res = oracle.full_batch_step(p, torch.from_numpy(x), a_hat_cpu, r_cpu, abnormal, normal,
                                         torch.randn(len(abnormal), h) * var + mean)
            res["loss"].backward()
            opt_c.step()
            tt.append(time.perf_counter() - t0)
        cpu_s = float(np.median(tt))
    out = {"config": args.config, "nodes": n, "nnz_A_hat": int(g_hat.nnz), "d": d, "h": h,
           "n_normal": len(normal), "n_outlier_seeds": len(abnormal),
           "gpu_epoch_ms_events": gpu_ms, "gpu_epoch_ms_wall": wall_ms, "ggad_launches_per_epoch": launches,
           "cuda_graph_epoch_ms_events": graph_ms, "cuda_graph_epoch_ms_wall": graph_wall_ms,
           "cpu_csr_oracle_epoch_s": cpu_s, "cpu_threads": torch.get_num_threads(),
           "speedup_vs_cpu_csr_oracle": (cpu_s * 1e3 / graph_wall_ms) if cpu_s else None,
           "reference_dense_epoch_s_photo_8vcpu_survey": 5.46 if args.config == "C1" else None,
           "loss": float(loss)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
