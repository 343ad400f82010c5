#!/usr/bin/env python
"""f3: evaluation of program B on the DGraph-shaped graph -- all test nodes scored in one layer-wise device pass
(ggad_b200.evaluate.to_prob_all / test_sage) against the reference's schedule (GCN.to_prob on batch_size-node batches,
src/utils.py:215-224), both on the GPU drop-in.  One JSON line.
    python tools/bench_eval.py [--test-nodes 1000000] [--batch 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=3_700_550)
ap.add_argument("--edges", type=int, default=36_552_754)
ap.add_argument("--test-nodes", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=200)
ap.add_argument("--loop-batches", type=int, default=100, help="batches of the reference schedule that are timed (extrapolated)")
a = ap.parse_args()
from ggad_b200 import evaluate, graphsage as gs, synth  # noqa: E402

dev = torch.device("cuda")
adj = synth.rmat_adjacency(a.nodes, a.edges, seed=72, device=dev)
rng = np.random.default_rng(72)
x = rng.random((a.nodes, 17), dtype=np.float32)
feats = torch.nn.Embedding(a.nodes, 17)
feats.weight = torch.nn.Parameter(torch.from_numpy(x), requires_grad=False)
feats = feats.to(dev)
torch.manual_seed(72)
enc = gs.GCNEncoder(feats, 17, 64, adj, gs.GCNAggregator(feats, cuda=True), gcn=True, cuda=True)
model = gs.GCN(2, enc).to(dev)
test = rng.permutation(a.nodes)[: a.test_nodes]
labels = (rng.random(len(test)) < 0.013).astype(np.int64)
evaluate.to_prob_all(model, test[:10000], a.batch)                      # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
res = evaluate.test_sage(test, labels, model, a.batch, verbose=False)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
nb = a.loop_batches
with torch.no_grad():
    model.to_prob(test[: a.batch].tolist(), None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(nb):
        model.to_prob(test[i * a.batch:(i + 1) * a.batch].tolist(), None)
    torch.cuda.synchronize()
t_loop = (time.perf_counter() - t0) / nb
print(json.dumps({"workload": "C4 evaluation (test_sage)", "test_nodes": len(test), "batch_size": a.batch,
                  "layerwise_one_pass_s": t_all, "nodes_per_s": len(test) / t_all,
                  "batched_loop_ms_per_batch": t_loop * 1e3, "batched_loop_s_extrapolated": t_loop * (len(test) / a.batch),
                  "speedup": t_loop * (len(test) / a.batch) / t_all, "auc": res[3]}))
