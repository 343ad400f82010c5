"""On-device evaluation of program B: the reference's ``test_sage`` (src/utils.py:207-247) scores
the test nodes in ``batch_size``-node batches through ``GCN.to_prob`` -- one Python round trip, one dense mask and
one sklearn call per batch / per evaluation.  Here ALL test nodes are scored in one layer-wise pass on the device:

  * the aggregation block of every test node is built at once (ggad_block_rowptr / ggad_block_fill);
  * the reference's *batch-local* column degree (src/graphsage.py:315: how many rows of the SAME batch contain
    node u) is kept exactly by de-duplicating the composite key (batch id, u) instead of u -- so the numbers equal
    the batched loop's, not a differently normalised full-graph variant;
  * one gather-reduce launch straight from the feature table (xmap), one projection (ggad_dense_matmul), sigmoid;
  * AUROC / AP / F1 / G-mean from the device-resident scores (metrics.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import metrics, ops
from ._lib import check, lib, ptr, stream_ptr
from .graph import AdjListCSR, CSRGraph


def to_prob_all(model, nodes: Sequence[int], batch_size: int) -> torch.Tensor:
    """sigmoid scores [T] of ``model`` (graphsage.GCN) for all ``nodes``, equal to concatenating
    ``model.to_prob(nodes[i:i+batch_size], None)`` over the batches (src/utils.py:215-224)."""
    from .graphsage import _device_of, _feature_table
    enc = model.enc
    dev = _device_of(enc.features)
    adj = enc.adj_lists if hasattr(enc.adj_lists, "block") else AdjListCSR.get(enc.adj_lists).device(dev)
    t = len(nodes)
    if t == 0:
        return torch.zeros(0, device=dev)
    nodes_d = (nodes if isinstance(nodes, torch.Tensor) else torch.as_tensor(np.asarray(nodes, dtype=np.int64))).to(dev, torch.int32)
    h = lib()
    with torch.cuda.device(dev), torch.no_grad():
        st = stream_ptr(dev)
        rowptr = torch.empty(t + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        check(h.ggad_block_rowptr(ptr(adj.rowptr), ptr(adj.col), adj.n, ptr(nodes_d), t, 1, ptr(rowptr), C.addressof(nnz), st))
        m = int(nnz.value)
        cols = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
        check(h.ggad_block_fill(ptr(adj.rowptr), ptr(adj.col), adj.n, ptr(nodes_d), t, 1, ptr(rowptr), ptr(cols), st))
        cols = cols[:m]
        rdeg = rowptr[1:] - rowptr[:-1]                                          # exact ints: |N(b) U {b}|
        batch_of_row = torch.arange(t, device=dev) // int(batch_size)
        key = torch.repeat_interleave(batch_of_row, rdeg) * int(adj.n) + cols.long()     # (batch, node) composite
        uniq, local, cdeg = torch.unique(key, return_inverse=True, return_counts=True)   # batch-local column degree
        g = CSRGraph(rowptr, local.to(torch.int32), None, t, int(uniq.numel()),
                     row_scale=1.0 / rdeg.to(torch.float32).sqrt(), col_scale=1.0 / cdeg.to(torch.float32).sqrt())
        table = _feature_table(enc.features)
        d = enc.features.weight.shape[1]
        to_feats = ops.gather_reduce(g, table, xmap=(uniq % int(adj.n)).to(torch.int32))["y"]   # sym-norm aggregate (:318-326)
        combined = ops.dense_matmul(to_feats, ops.pad_cols(enc.weight.detach()), trans_b=True, relu=True)  # ReLU(W agg^T) (:412)
        scores = ops.dense_matmul(combined, model.weight.detach(), trans_b=True)          # weight . embeds (:174)
        return torch.sigmoid(scores[:, 0])


def test_sage(test_cases: Sequence[int], labels, model, batch_size: int, thres: float = 0.5, verbose: bool = True):
    """Drop-in for the reference's ``test_sage`` (same arguments and return tuple
    ``(f1_macro, f1_binary_1, f1_binary_0, auc, gmean)``); everything up to the five scalars stays on the device."""
    prob = to_prob_all(model, test_cases, batch_size)
    y = torch.as_tensor(np.asarray(labels)).to(prob.device).reshape(-1)
    pred = (prob >= thres).to(torch.int64)                       # prob2pred (src/utils.py:250-256)
    yi = y.to(torch.int64)
    tp = ((pred == 1) & (yi == 1)).sum().double()
    tn = ((pred == 0) & (yi == 0)).sum().double()
    fp = ((pred == 1) & (yi == 0)).sum().double()
    fn = ((pred == 0) & (yi == 1)).sum().double()

    def f1(t_, f_p, f_n):
        den = 2 * t_ + f_p + f_n
        return torch.where(den > 0, 2 * t_ / den.clamp_min(1), torch.zeros_like(den))
    f1_1, f1_0 = f1(tp, fp, fn), f1(tn, fn, fp)
    auc = metrics.roc_auc(prob, y)
    ap = metrics.average_precision(prob, y)
    gmean = torch.sqrt((tp / (tp + fn).clamp_min(1)) * (tn / (tn + fp).clamp_min(1)))   # conf_gmean (src/utils.py:324-326)
    out = [float(v) for v in torch.stack([(f1_1 + f1_0) / 2, f1_1, f1_0, auc, gmean, ap, tp, tn, fn, fp]).cpu()]
    if verbose:
        print(f"   GNN F1-binary-1: {out[1]:.4f}\tF1-binary-0: {out[2]:.4f}\tF1-macro: {out[0]:.4f}\tG-Mean: {out[4]:.4f}\tAUC: {out[3]:.4f}")
        print('Testing AP:', out[5])
        print(f"   GNN TP: {int(out[6])}\tTN: {int(out[7])}\tFN: {int(out[8])}\tFP: {int(out[9])}")
    return out[0], out[1], out[2], out[3], out[4]
