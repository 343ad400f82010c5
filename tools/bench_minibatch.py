#!/usr/bin/env python
"""Program B (mini-batch GGAD on a DGraph-shaped graph): one training batch = GCN.loss + backward + Adam
(src/model_handler.py:330-364) with the drop-in modules, frontier built on the device.

    python tools/bench_minibatch.py [--nodes 3700550 --edges 36552754 --batch 150 --seeds 50] [--cpu-nodes 300000]

Also times the CPU restatement of the reference's batch (oracle, sparse; the reference's own dense-mask batch
took 1.52 s on a 300 k-node proxy in the survey) on a smaller graph of the same generator.  One JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=3_700_550)
    ap.add_argument("--edges", type=int, default=36_552_754, help="directed R-MAT candidates (symmetrised afterwards)")
    ap.add_argument("--d", type=int, default=17)
    ap.add_argument("--h", type=int, default=64)
    ap.add_argument("--batch", type=int, default=150)
    ap.add_argument("--seeds", type=int, default=50)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--cpu-nodes", type=int, default=300_000)
    ap.add_argument("--cpu-iters", type=int, default=3)
    args = ap.parse_args()
    from ggad_b200 import _lib, graphsage as gs, synth
    dev = torch.device("cuda")
    adj = synth.rmat_adjacency(args.nodes, args.edges, seed=72, device=dev)
    n, d, h = args.nodes, args.d, args.h
    rng = np.random.default_rng(72)
    x = rng.random((n, d), dtype=np.float32)
    x = x / (x.sum(1, keepdims=True) + 0.01)                      # src/utils.py:79 normalisation
    feats = torch.nn.Embedding(n, d)
    feats.weight = torch.nn.Parameter(torch.from_numpy(x), requires_grad=False)
    feats = feats.cuda()
    agg = gs.GCNAggregator(feats, cuda=True)
    enc = gs.GCNEncoder(feats, d, h, adj, agg, gcn=True, cuda=True)
    model = gs.GCN(2, enc).cuda()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.007)
    deg = (adj.rowptr[1:] - adj.rowptr[:-1]).cpu().numpy()
    cand = np.flatnonzero(deg > 0)
    B = args.batch + args.seeds
    labels = torch.cat([torch.zeros(args.batch, dtype=torch.long), torch.ones(args.seeds, dtype=torch.long)])

    def batch(i):
        nodes = rng.choice(cand, B, replace=False).tolist()
        opt.zero_grad()
        total, cls, margin, rec = model.loss(nodes, labels)
        total.backward()
        opt.step()
        return total

    stats = []
    for i in range(args.iters + 3):
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        loss = batch(i)
        lv = float(loss.detach())
        torch.cuda.synchronize()
        hop1, hop2 = agg.last_blocks
        stats.append(((time.perf_counter() - t0) * 1e3, _lib.launch_count() - l0, hop1.n_cols,
                      int(hop1.col_d.numel()), hop2.n_cols, int(hop2.col_d.numel())))
    st = np.array(stats[3:], dtype=np.float64)
    ms = float(np.median(st[:, 0]))
    edges = float(np.mean(st[:, 3] + st[:, 5]))
    out = {"workload": "C4 mini-batch GGAD", "nodes": n, "adjacency_entries": int(adj.col.numel()), "d": d, "h": h,
           "batch": B, "ms_per_batch": ms, "batches_per_s": 1e3 / ms, "ggad_launches_per_batch": float(np.mean(st[:, 1])),
           "mean_frontier_U": float(np.mean(st[:, 2])), "mean_hop1_edges": float(np.mean(st[:, 3])),
           "mean_frontier_U2": float(np.mean(st[:, 4])), "mean_hop2_edges": float(np.mean(st[:, 5])),
           "edges_per_s": edges / (ms * 1e-3), "loss": lv,
           "reference_cpu_s_per_batch_300k_proxy_survey": 1.52}
    if args.cpu_nodes > 0:
        import oracle
        na = args.cpu_nodes
        adj_c = synth.rmat_adjacency(na, int(args.edges * na / n), seed=72, device=dev)
        rp, col = adj_c.rowptr.cpu().numpy(), adj_c.col.cpu().numpy()
        t0 = time.perf_counter()
        adj_lists = {i: set(col[rp[i]:rp[i + 1]].tolist()) for i in range(na)}
        t_dict = time.perf_counter() - t0
        xc = torch.from_numpy(x[:na].copy())
        p = {"enc.weight": enc.weight.detach().cpu().clone().requires_grad_(True),
             "enc.fc.weight": enc.fc.weight.detach().cpu().clone().requires_grad_(True),
             "weight": model.weight.detach().cpu().clone().requires_grad_(True)}
        degc = np.diff(rp)
        candc = np.flatnonzero(degc > 0)
        tt = []
        for _ in range(args.cpu_iters):
            nodes = rng.choice(candc, B, replace=False).tolist()
            t0 = time.perf_counter()
            total = oracle.gcn_minibatch_loss(p, nodes, labels, adj_lists, xc)[0]
            total.backward()
            tt.append(time.perf_counter() - t0)
        out.update({"cpu_oracle_s_per_batch": float(np.median(tt)), "cpu_graph_nodes": na, "cpu_threads": torch.get_num_threads(),
                    "cpu_dict_build_s": t_dict})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
