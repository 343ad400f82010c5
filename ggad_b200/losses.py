"""The loss block of the reference's full-batch training loop (run.py:164-210) on CSR.

run.py is a script, so this is the one piece of program A a caller has to import instead of
inlining: ``ggad_loss`` returns the same scalars run.py prints, with the N x N similarity matrix
replaced by the row-subset local-affinity kernel (only aff[normal] and aff[abnormal] are consumed).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .graph import CSRGraph

_subset_cache = {}


def _subset(normal_idx, abnormal_idx, device):
    """Unique union of the two index lists + positions of each list inside it.  Cached per list objects (the
    cache holds references, so ids cannot be recycled under it)."""
    key = (id(normal_idx), id(abnormal_idx), str(device))
    hit = _subset_cache.get(key)
    n, a = np.asarray(normal_idx, dtype=np.int64), np.asarray(abnormal_idx, dtype=np.int64)
    probe = (hash(n.tobytes()), hash(a.tobytes()))            # content key: in-place edits of the lists are seen
    if hit is None or hit[0] is not normal_idx or hit[1] is not abnormal_idx or hit[2] != probe:
        uniq, inv = np.unique(np.concatenate([n, a]), return_inverse=True)
        val = (torch.from_numpy(uniq.astype(np.int32)).to(device),
               torch.from_numpy(inv[:len(n)]).to(device), torch.from_numpy(inv[len(n):]).to(device))
        hit = (normal_idx, abnormal_idx, probe, val)
        if len(_subset_cache) > 8:
            _subset_cache.clear()
        _subset_cache[key] = hit
    return hit[3]


def ggad_loss(emb, logits, emb_con, emb_abnormal, raw_adj, normal_label_idx, abnormal_label_idx,
              negsamp_ratio: float = 1.0, confidence_margin: float = 0.7):
    """(loss, loss_margin, loss_bce, loss_rec, affinity_normal_mean, affinity_abnormal_mean).

    emb [1,N,h] is Model.forward's first output (after the write-back), raw_adj is R = A + I as a
    CSRGraph (or anything CSRGraph.from_any accepts)."""
    device = emb.device
    from .model import as_graph
    g_r = as_graph(raw_adj, device)            # converted once and cached (keyed on the object and its _version)
    # BCE (run.py:165-172): labels are [0]*|normal| + [1]*|S|
    lbl = torch.cat((torch.zeros(len(normal_label_idx), device=device),
                     torch.ones(emb_con.shape[0], device=device))).unsqueeze(1).unsqueeze(0)
    loss_bce = F.binary_cross_entropy_with_logits(logits, lbl, reduction='none',
                                                  pos_weight=torch.full((1,), float(negsamp_ratio), device=device)).mean()
    # local affinity (run.py:175-191) on the consumed rows only
    e = emb[0] if emb.dim() == 3 else emb
    subset, pos_n, pos_a = _subset(normal_label_idx, abnormal_label_idx, device)
    aff = ops.local_affinity(e, g_r, subset)
    aff_n, aff_a = aff[pos_n].mean(), aff[pos_a].mean()
    loss_margin = (confidence_margin - (aff_n - aff_a)).clamp_min(0)
    # run.py:207-208 -- emb_abnormal carries the batch dim, so the sum runs over the |S| axis
    diff = torch.pow(emb_con - emb_abnormal, 2)
    loss_rec = torch.mean(torch.sqrt(torch.sum(diff, 1)))
    loss = 1 * loss_margin + 1 * loss_bce + 1 * loss_rec
    return loss, loss_margin, loss_bce, loss_rec, aff_n, aff_a


# ---------------------------------------------------------------------------------------------------------------
# the same local-affinity path as the other full-batch script of the reference uses it (tam.py:113-146)
# ---------------------------------------------------------------------------------------------------------------
def _row_col_scale(g: CSRGraph):
    """rowsum_i / colsum_i of the stored values (0 where the column sum is 0), cached on the graph.  tam.py divides the
    ROW sums of sim * adj by the COLUMN sums of adj; ops.local_affinity divides by the sums of the rows it reduces."""
    hit = g.__dict__.get("_row_col_scale")
    if hit is None:
        def row_sums(c):
            ones = torch.ones(c.n_cols, 4, dtype=torch.float32, device=c.device)
            return ops.gather_reduce(c, ones, use_graph_scales=False)["y"][:, 0]
        rs, cs = row_sums(g), row_sums(g.T)
        hit = torch.where(cs != 0, rs / cs, torch.zeros_like(cs))
        g.__dict__["_row_col_scale"] = hit
    return hit


def inference(feature, adj_matrix):
    """``tam.inference`` (``tam.py:136-146``): message_i = sum_j adj[i,j] <f^_i, f^_j> / sum_k adj[k,i] for every node,
    without the N x N similarity matrix.  ``adj_matrix`` is anything ``model.as_graph`` accepts (dense [N,N] tensor, scipy,
    CSRGraph).  A node with an all-zero feature row contributes 0 (the convention of ``max_message`` and of run.py:177-180;
    the reference's ``inference`` would return NaN there)."""
    from .model import as_graph
    f = feature[0] if feature.dim() == 3 else feature
    g = as_graph(adj_matrix, f.device)
    rows = g.__dict__.get("_all_rows")
    if rows is None:
        rows = g.__dict__["_all_rows"] = torch.arange(g.n_rows, dtype=torch.int32, device=f.device)
    # local_affinity(e, R, S)_j = sum_i R[i,j] <e^_i, e^_j> / sum_i R[i,j]; with R = adj^T that is the row-wise sum over adj
    return ops.local_affinity(f, g.T, rows) * _row_col_scale(g)


def max_message(feature, adj_matrix, normal_label_idx):
    """``tam.max_message`` (``tam.py:113-133``): the messages of ``inference`` min-max normalised over all nodes;
    returns ``(-sum(message[normal_label_idx]), message)``."""
    m = inference(feature, adj_matrix)
    lo, hi = m.min(), m.max()
    m = (m - lo) / (hi - lo)
    idx = torch.as_tensor(np.asarray(normal_label_idx, dtype=np.int64), device=m.device)
    return -m[idx].sum(), m
