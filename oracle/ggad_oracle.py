"""CPU restatement of the GGAD hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Everything here is O(nnz) CSR / COO arithmetic in numpy, scipy and torch-CPU
fp32 (degree normalisation in fp64 exactly as the reference does), written from
the reference's *behaviour*; every function cites the reference lines it
restates (paths are relative to /root/reference).

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz, which were
produced by importing the reference's own ``model.py`` / ``src/graphsage.py``
(see tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
import random as _random
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

__all__ = [
    "normalize_adj", "preprocess_features", "build_full_batch_graph", "csr_arrays",
    "spmm_csr", "spmm_csr_rows", "gcn_layer", "model_forward", "local_affinity",
    "full_batch_losses", "full_batch_step", "load_mat_split", "normalize_rows_minibatch",
    "adj_lists_to_csr", "gcn_aggregator", "gcn_encoder", "gcn_minibatch_forward",
    "gcn_minibatch_loss", "mean_aggregator", "sage_encoder", "torch_spmm_cpu_baseline",
    "torch_spmm_cpu_operator", "torch_spmm_cpu_step",
]


# --------------------------------------------------------------------------
# L0: preprocessing (utils.py)
# --------------------------------------------------------------------------
def normalize_adj(adj) -> sp.coo_matrix:
    """``D^-1/2 . A^T . D^-1/2`` in fp64 with D from ROW sums (utils.py:47-54).

    The product is formed as ``(A.D)^T . D`` so entry [i,j] = A[j,i]*d_i*d_j with
    the two fp64 multiplications in that order; 1/sqrt(0) is mapped to 0.
    """
    a = sp.coo_matrix(adj)
    deg = np.asarray(a.sum(1), dtype=np.float64).reshape(-1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)
    dis[np.isinf(dis)] = 0.0
    dmat = sp.diags(dis)
    return a.dot(dmat).transpose().dot(dmat).tocoo()


def preprocess_features(features) -> np.ndarray:
    """Row-normalise a feature matrix, 1/0 -> 0 (utils.py:37-44)."""
    f = sp.csr_matrix(features)
    rs = np.asarray(f.sum(1), dtype=np.float64).reshape(-1)
    with np.errstate(divide="ignore"):
        inv = np.power(rs, -1.0)
    inv[np.isinf(inv)] = 0.0
    return np.asarray(sp.diags(inv).dot(f).todense())


def build_full_batch_graph(adj) -> Tuple[sp.csr_matrix, sp.csr_matrix]:
    """(A_hat, R) exactly as run.py:96-109 builds ``adj`` and ``raw_adj``.

    A_hat = normalize_adj(A) + I   (self loop added AFTER normalisation, weight 1)
    R     = A + I                  (un-normalised, weighted)
    Both are formed in fp64 and rounded once to fp32 (the reference densifies the
    fp64 matrix and wraps it in torch.FloatTensor).
    """
    a = sp.csr_matrix(adj).astype(np.float64)
    n = a.shape[0]
    eye = sp.eye(n, dtype=np.float64, format="csr")
    a_hat = (sp.csr_matrix(normalize_adj(a)) + eye).tocsr()
    r = (a + eye).tocsr()
    for m in (a_hat, r):
        m.sum_duplicates()
        m.sort_indices()
    return a_hat.astype(np.float32), r.astype(np.float32)


def csr_arrays(m: sp.csr_matrix) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    m = sp.csr_matrix(m)
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float32)


# --------------------------------------------------------------------------
# L1: neighbor gather-reduce
# --------------------------------------------------------------------------
def _row_ids(rowptr: np.ndarray) -> np.ndarray:
    return np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))


def spmm_csr(rowptr, col, val, x: torch.Tensor, n_rows: Optional[int] = None) -> torch.Tensor:
    """Y[r] = sum_e val[e] * X[col[e]] over the CSR row r  (what ``torch.spmm`` /
    ``torch.bmm`` with the dense adjacency compute at model.py:29,31).  fp32,
    sequential accumulation in CSR order; differentiable w.r.t. ``x``."""
    rowptr = np.asarray(rowptr)
    n = len(rowptr) - 1 if n_rows is None else n_rows
    rows = torch.from_numpy(_row_ids(rowptr))
    colt = torch.from_numpy(np.asarray(col).astype(np.int64))
    msg = x[colt]
    if val is not None:
        msg = msg * torch.as_tensor(np.asarray(val), dtype=x.dtype).unsqueeze(1)
    out = torch.zeros(n, x.shape[1], dtype=x.dtype)
    return out.index_add(0, rows, msg)


def spmm_csr_rows(rowptr, col, val, x: torch.Tensor, rows: Sequence[int]) -> torch.Tensor:
    """Rows ``rows`` of A.X -- the ego-neighbor weighted sum ``adj[0,S,:] @ emb``
    of model.py:151-155 without densifying."""
    rowptr = np.asarray(rowptr)
    rows = np.asarray(rows, dtype=np.int64)
    lens = rowptr[rows + 1] - rowptr[rows]
    sub_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_ptr[1:])
    eidx = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    eidx = eidx.astype(np.int64)
    v = None if val is None else np.asarray(val)[eidx]
    return spmm_csr(sub_ptr, np.asarray(col)[eidx], v, x, n_rows=len(rows))


def gcn_layer(x: torch.Tensor, a_hat, weight: torch.Tensor, bias: Optional[torch.Tensor],
              prelu: Optional[torch.Tensor]) -> torch.Tensor:
    """One reference GCN layer: project, aggregate, add bias, PReLU (model.py:26-35)."""
    rowptr, col, val = a_hat
    out = spmm_csr(rowptr, col, val, x @ weight.t())
    if bias is not None:
        out = out + bias
    if prelu is not None:
        out = torch.nn.functional.prelu(out, prelu.reshape(-1))   # nn.PReLU (model.py:10,35): slope also at exactly 0
    return out


def model_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, a_hat, sample_abnormal_idx: Sequence[int],
                  normal_idx: Sequence[int], train_flag: bool, noise: torch.Tensor):
    """Reference ``Model.forward`` (model.py:133-191) on a CSR adjacency.

    ``p`` uses the reference state_dict keys.  ``noise`` is the explicit tensor the
    reference draws as ``randn*var+mean`` at model.py:143.  Returns the same
    5-tuple *without* the leading batch dimension: (emb[N,h] AFTER the in-place
    write-back, emb_combine, f_3, emb_con, emb_abnormal).
    """
    h1 = gcn_layer(x, a_hat, p["gcn1.fc.weight"], p["gcn1.bias"], p["gcn1.act.weight"])
    emb = gcn_layer(h1, a_hat, p["gcn2.fc.weight"], p["gcn2.bias"], p["gcn2.act.weight"])
    s = torch.as_tensor(np.asarray(sample_abnormal_idx, dtype=np.int64))
    nrm = torch.as_tensor(np.asarray(normal_idx, dtype=np.int64))
    emb_abnormal = emb[s] + noise                                        # model.py:141-144

    def mlp(t):                                                          # model.py:176-180
        t = F.relu(t @ p["fc1.weight"].t())
        t = F.relu(t @ p["fc2.weight"].t())
        return t @ p["fc3.weight"].t()

    if not train_flag:                                                   # model.py:183-188
        return emb, None, mlp(emb), None, emb_abnormal
    rowptr, col, val = a_hat
    ego = spmm_csr_rows(rowptr, col, val, emb, sample_abnormal_idx)      # model.py:151-155
    emb_con = F.relu(ego @ p["fc4.weight"].t())                          # model.py:156
    emb_combine = torch.cat((emb[nrm], emb_con), 0)                      # model.py:159
    f3 = mlp(emb_combine)
    emb_out = emb.index_copy(0, s, emb_con)                              # model.py:182 (in-place there)
    return emb_out, emb_combine, f3, emb_con, emb_abnormal


def local_affinity(emb: torch.Tensor, r_csr) -> torch.Tensor:
    """aff_j = sum_i R[i,j] <e^_i, e^_j> / sum_i R[i,j]   (run.py:175-188).

    Column (axis-0) reductions of ``sim * raw_adj``; 1/||e|| and 1/colsum map
    inf -> 0.  ``r_csr`` is the CSR of R=A+I (rows i, cols j).
    """
    rowptr, col, val = r_csr
    n = emb.shape[0]
    nrm = torch.norm(emb, dim=-1, keepdim=True)
    inv = torch.pow(nrm, -1)
    inv = torch.where(torch.isinf(inv), torch.zeros_like(inv), inv)
    e = emb * inv
    i = torch.from_numpy(_row_ids(np.asarray(rowptr)))
    j = torch.from_numpy(np.asarray(col).astype(np.int64))
    w = torch.as_tensor(np.asarray(val), dtype=emb.dtype)
    dots = (e[i] * e[j]).sum(1) * w
    num = torch.zeros(n, dtype=emb.dtype).index_add(0, j, dots)
    den = torch.zeros(n, dtype=emb.dtype).index_add(0, j, w)
    r_inv = torch.pow(den, -1)
    r_inv = torch.where(torch.isinf(r_inv), torch.zeros_like(r_inv), r_inv)
    return num * r_inv


def full_batch_losses(emb, logits, emb_con, emb_abnormal, r_csr, normal_idx, abnormal_idx,
                      pos_weight: float = 1.0, margin_c: float = 0.7):
    """The loss block of run.py:164-210: (loss, margin, bce, rec, affinity)."""
    n_norm, n_abn = len(normal_idx), emb_con.shape[0]
    lbl = torch.cat((torch.zeros(n_norm), torch.ones(n_abn))).unsqueeze(1)
    bce = F.binary_cross_entropy_with_logits(logits, lbl, reduction="none",
                                             pos_weight=torch.tensor([float(pos_weight)])).mean()
    aff = local_affinity(emb, r_csr)
    nrm = torch.as_tensor(np.asarray(normal_idx, dtype=np.int64))
    abn = torch.as_tensor(np.asarray(abnormal_idx, dtype=np.int64))
    margin = (margin_c - (aff[nrm].mean() - aff[abn].mean())).clamp_min(0)
    # run.py:207-208: emb_abnormal carries the batch dim ([1,|S|,h]) so the
    # broadcast difference is [1,|S|,h] and ``torch.sum(.., 1)`` reduces over the
    # |S| axis (NOT over h): rec = mean_k sqrt(sum_j (con[j,k]-abn[j,k])^2).
    rec = torch.sqrt(torch.sum((emb_con - emb_abnormal) ** 2, 0)).mean()
    return margin + bce + rec, margin, bce, rec, aff


def full_batch_step(p, x, a_hat, r_csr, abnormal_idx, normal_idx, noise):
    """One training forward of program A (run.py:152-210) -> dict of tensors."""
    emb, comb, f3, emb_con, emb_abn = model_forward(p, x, a_hat, abnormal_idx, normal_idx, True, noise)
    loss, margin, bce, rec, aff = full_batch_losses(emb, f3, emb_con, emb_abn, r_csr, normal_idx, abnormal_idx)
    return dict(emb=emb, emb_combine=comb, logits=f3, emb_con=emb_con, emb_abnormal=emb_abn,
                loss=loss, margin=margin, bce=bce, rec=rec, affinity=aff)


def load_mat_split(ano_labels: np.ndarray, dataset: str, seed: int, train_rate=0.3, val_rate=0.1):
    """Semi-supervised split of utils.py:89-141 driven by Python's ``random``."""
    rng = _random.Random(seed)
    n = len(ano_labels)
    all_idx = list(range(n))
    rng.shuffle(all_idx)
    n_tr, n_va = int(n * train_rate), int(n * val_rate)
    idx_train, idx_val, idx_test = all_idx[:n_tr], all_idx[n_tr:n_tr + n_va], all_idx[n_tr + n_va:]
    normal = [i for i in idx_train if ano_labels[i] == 0]
    normal = normal[: int(len(normal) * 0.5)]
    rng.shuffle(normal)
    frac = 0.05 if dataset == "Amazon" else 0.15
    abnormal = normal[: int(len(normal) * frac)]
    return idx_train, idx_val, idx_test, normal, abnormal


# --------------------------------------------------------------------------
# Program B: mini-batch GGAD on adjacency lists (src/graphsage.py)
# --------------------------------------------------------------------------
def normalize_rows_minibatch(mx: np.ndarray) -> np.ndarray:
    """``(rowsum + 0.01)^-1`` feature scaling of src/utils.py:74-84."""
    rs = np.asarray(mx.sum(1), dtype=np.float64).reshape(-1) + 0.01
    inv = np.power(rs, -1.0)
    inv[np.isinf(inv)] = 0.0
    return np.asarray(sp.diags(inv).dot(mx))


def adj_lists_to_csr(adj_lists: Dict[int, Iterable[int]], n: int):
    """dict[int -> set[int]] (the un-pickled format of src/utils.py:27-28,96-112)
    -> CSR with sorted columns, no values."""
    rowptr = np.zeros(n + 1, dtype=np.int64)
    for k, v in adj_lists.items():
        rowptr[int(k) + 1] = len(v)
    np.cumsum(rowptr, out=rowptr)
    col = np.zeros(rowptr[-1], dtype=np.int32)
    for k, v in adj_lists.items():
        k = int(k)
        col[rowptr[k]:rowptr[k + 1]] = sorted(int(t) for t in v)
    return rowptr, col


def _block(neigh_sets: List[List[int]]):
    """Union frontier (sorted) and the 0/1 block in COO form, with exact int degrees."""
    frontier = sorted(set().union(*[set(s) for s in neigh_sets])) if neigh_sets else []
    pos = {n: i for i, n in enumerate(frontier)}
    rows = np.fromiter((i for i, s in enumerate(neigh_sets) for _ in s), dtype=np.int64)
    cols = np.fromiter((pos[n] for s in neigh_sets for n in s), dtype=np.int64)
    rdeg = np.array([len(s) for s in neigh_sets], dtype=np.int64)
    cdeg = np.bincount(cols, minlength=len(frontier)).astype(np.int64)
    return frontier, rows, cols, rdeg, cdeg


def _sym_weights(rows, cols, rdeg, cdeg) -> torch.Tensor:
    """mask.div(sqrt(rowsum)).div(sqrt(colsum)) in fp32 (src/graphsage.py:314-318)."""
    r = torch.as_tensor(rdeg, dtype=torch.float32).sqrt()
    c = torch.as_tensor(cdeg, dtype=torch.float32).sqrt()
    one = torch.ones(len(rows), dtype=torch.float32)
    return one.div(r[torch.from_numpy(rows)]).div(c[torch.from_numpy(cols)])


def gcn_aggregator(nodes: Sequence[int], adj_lists, feats: torch.Tensor, train_flag: bool):
    """``GCNAggregator.forward`` CPU branch (src/graphsage.py:295-360).

    Returns dict with to_feats[B,d], to_feats_neigh[|U|,d] or None, the hop-1
    frontier U (sorted; the reference uses Python-set order, results are
    permutation-equivalent), and the mean mask ``mask_row`` as COO (rows, cols,
    1/rdeg).  Hop-1 unions self, hop-2 does not; column degrees are batch-local.
    Empty hop-2 rows give 0/0 = NaN exactly as the dense reference does.
    """
    nodes = [int(n) for n in nodes]
    s1 = [sorted(set(adj_lists[n]) | {n}) for n in nodes]
    u, r1, c1, rdeg1, cdeg1 = _block(s1)
    w1 = _sym_weights(r1, c1, rdeg1, cdeg1)
    ut = torch.as_tensor(np.asarray(u, dtype=np.int64))
    xu = feats[ut]
    to_feats = torch.zeros(len(nodes), feats.shape[1]).index_add(0, torch.from_numpy(r1), xu[torch.from_numpy(c1)] * w1[:, None])
    out = dict(to_feats=to_feats, U=u, rows=r1, cols=c1, rdeg=rdeg1, cdeg=cdeg1,
               mask_row_w=torch.ones(len(r1)).div(torch.as_tensor(rdeg1, dtype=torch.float32)[torch.from_numpy(r1)]),
               to_feats_neigh=None)
    if train_flag:
        s2 = [sorted(set(adj_lists.get(n))) for n in u]
        u2, r2, c2, rdeg2, cdeg2 = _block(s2)
        w2 = _sym_weights(r2, c2, rdeg2, cdeg2)
        xu2 = feats[torch.as_tensor(np.asarray(u2, dtype=np.int64))] if len(u2) else feats[:0]
        tfn = torch.zeros(len(u), feats.shape[1]).index_add(0, torch.from_numpy(r2), xu2[torch.from_numpy(c2)] * w2[:, None])
        empty = torch.as_tensor(rdeg2 == 0)
        if bool(empty.any()):
            tfn = torch.where(empty[:, None], torch.full_like(tfn, float("nan")), tfn)
        out.update(to_feats_neigh=tfn, U2=u2, rows2=r2, cols2=c2, rdeg2=rdeg2, cdeg2=cdeg2)
    return out


def gcn_encoder(p, nodes, labels: torch.Tensor, adj_lists, feats, train_flag: bool):
    """``GCNEncoder.forward`` (src/graphsage.py:395-454).  ``p`` holds ``enc.weight``
    [h,d] and ``enc.fc.weight`` [h,h].  Returns (combined_all[h,B'], ego[B,h] or
    None, anomaly_feat[h,n1], anomaly_feat_new[h,n1])."""
    agg = gcn_aggregator(nodes, adj_lists, feats, train_flag)
    w = p["enc.weight"]
    combined = F.relu(w.mm(agg["to_feats"].t()))
    if not train_flag:
        return combined, None, None, None
    emb_u = F.relu(w.mm(agg["to_feats_neigh"].t()))                      # [h,|U|]
    r, c = torch.from_numpy(agg["rows"]), torch.from_numpy(agg["cols"])
    ego = torch.zeros(len(nodes), w.shape[0]).index_add(0, r, emb_u.t()[c] * agg["mask_row_w"][:, None])
    lab1, lab0 = labels == 1, labels == 0
    anomaly_feat = combined[:, lab1]
    anomaly_feat_new = F.relu(ego[lab1] @ p["enc.fc.weight"].t())
    combined_all = torch.cat((combined[:, lab0], anomaly_feat_new.t()), 1)   # label-0 columns first (:450)
    return combined_all, ego, anomaly_feat, anomaly_feat_new.t()


def gcn_minibatch_forward(p, nodes, labels, adj_lists, feats, train_flag):
    """``GCN.forward`` (src/graphsage.py:171-176): scores = (weight . embeds)^T."""
    embeds, ego, af, afn = gcn_encoder(p, nodes, labels, adj_lists, feats, train_flag)
    return p["weight"].mm(embeds).t(), ego, embeds, af, afn


def gcn_minibatch_loss(p, nodes, labels: torch.Tensor, adj_lists, feats):
    """``GCN.loss`` (src/graphsage.py:244-258) -> (total, cls, margin, rec)."""
    scores, ego, embeds, af, afn = gcn_minibatch_forward(p, nodes, labels, adj_lists, feats, True)
    cls = F.binary_cross_entropy_with_logits(scores.squeeze(), labels.float(), reduction="none",
                                             pos_weight=torch.tensor([1.0])).mean()
    aff = torch.cosine_similarity(embeds, ego.t(), dim=0)                # :234
    # :236-240 -- argwhere indexing keeps a trailing axis, so margin/total have shape [1]
    margin = (1 - (aff[torch.argwhere(labels == 0)].mean(0) - aff[torch.argwhere(labels == 1)].mean(0))).clamp_min(0)
    rec = torch.sqrt(torch.sum((af - afn) ** 2, 0)).mean()               # :197-198
    return cls + margin + 0.1 * rec, cls, margin, rec


def mean_aggregator(nodes, to_neighs: List[Iterable[int]], feats: torch.Tensor, gcn: bool = False):
    """``MeanAggregator.forward`` with ``num_sample=None`` (src/graphsage.py:66-99):
    per-row mean of neighbor features; isolated rows give 0/0 = NaN."""
    sets = [sorted(set(s) | ({int(nodes[i])} if gcn else set())) for i, s in enumerate(to_neighs)]
    u, r, c, rdeg, _ = _block(sets)
    w = torch.ones(len(r)).div(torch.as_tensor(rdeg, dtype=torch.float32)[torch.from_numpy(r)])
    xu = feats[torch.as_tensor(np.asarray(u, dtype=np.int64))] if len(u) else feats[:0]
    out = torch.zeros(len(sets), feats.shape[1]).index_add(0, torch.from_numpy(r), xu[torch.from_numpy(c)] * w[:, None])
    empty = torch.as_tensor(rdeg == 0)
    if bool(empty.any()):
        out = torch.where(empty[:, None], torch.full_like(out, float("nan")), out)
    return out


def sage_encoder(weight: torch.Tensor, nodes, adj_lists, feats: torch.Tensor, gcn: bool = False):
    """``Encoder.forward`` (src/graphsage.py:131-154): ReLU(W . cat(self, mean)^T)."""
    neigh = mean_aggregator(nodes, [adj_lists[int(n)] for n in nodes], feats, gcn=gcn)
    if gcn:
        comb = neigh
    else:
        comb = torch.cat((feats[torch.as_tensor(np.asarray(nodes, dtype=np.int64))], neigh), 1)
    return F.relu(weight.mm(comb.t()))


# --------------------------------------------------------------------------
# CPU baseline: the reference's own sparse branch (model.py:28-29 torch.spmm)
# --------------------------------------------------------------------------
def torch_spmm_cpu_operator(rowptr, col, val, n_cols: int):
    """The sparse ``adj`` the reference hands to torch.spmm (model.py:28-29) as a torch CSR tensor, built ONCE --
    run.py builds its adjacency once before the epoch loop, so timed steps must not pay for the conversion."""
    n = len(rowptr) - 1
    v = torch.ones(len(col)) if val is None else torch.as_tensor(np.asarray(val), dtype=torch.float32)
    return torch.sparse_csr_tensor(torch.as_tensor(np.asarray(rowptr, dtype=np.int64)),
                                   torch.as_tensor(np.asarray(col, dtype=np.int64)), v, size=(n, n_cols))


def torch_spmm_cpu_step(a, x: torch.Tensor, backward: bool = True):
    """One timed step of the reference's sparse branch: ``torch.spmm(adj, x)`` (model.py:29) on all host threads
    plus autograd's backward of loss = |y|^2 / 2.  Returns (y, dx or None)."""
    if not backward:
        return torch.spmm(a, x), None
    xr = x.detach().requires_grad_(True)
    y = torch.spmm(a, xr)
    y.backward(y.detach())
    return y.detach(), xr.grad


def torch_spmm_cpu_baseline(rowptr, col, val, x: torch.Tensor, backward: bool = True):
    """Operator construction + one step (kept for the small parity tests)."""
    return torch_spmm_cpu_step(torch_spmm_cpu_operator(rowptr, col, val, x.shape[0]), x, backward)
