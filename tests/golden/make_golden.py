#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE modules themselves.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):   python tests/golden/make_golden.py

* Program A: ``/root/reference/model.py`` (Model, GCN) is imported as-is and
  driven with dense adjacency exactly as run.py:96-109,152-154 does; the loss
  block run.py:164-210 is a script body (not importable), so it is evaluated
  here on the dense tensors the reference would hold, line for line in meaning.
* Program B: ``/root/reference/src/graphsage.py`` is imported with a stub for
  its unused ``torch_geometric`` import (src/graphsage.py:8).
* Data entry: ``/root/reference/utils.py`` is imported with stubs for ``dgl`` and
  ``matplotlib`` (used only by functions outside the path) and its ``load_mat``,
  ``preprocess_features`` and ``normalize_adj`` are run on synthetic ``.mat`` files
  (``data_*`` cases).

Each .npz stores inputs, the reference state_dict (``p/<key>``), outputs
(``o/<name>``) and parameter gradients (``g/<key>``).
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load_ref():
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF, "model.py"))
    ref_model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_model)
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.GCNConv = object
    tg.nn = tgnn
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.nn", tgnn)
    spec = importlib.util.spec_from_file_location("ref_graphsage", os.path.join(REF, "src", "graphsage.py"))
    ref_sage = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_sage)
    return ref_model, ref_sage


def _ref_normalize_adj(adj):
    # utils.py cannot be imported (dgl / networkx-2 imports); this is its
    # normalize_adj (utils.py:47-54) evaluated with the same scipy calls.
    adj = sp.coo_matrix(adj)
    rowsum = np.array(adj.sum(1))
    with np.errstate(divide="ignore"):
        d_inv_sqrt = np.power(rowsum, -0.5).flatten()
    d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.0
    d_mat = sp.diags(d_inv_sqrt)
    return adj.dot(d_mat).transpose().dot(d_mat).tocoo()


# ----------------------------------------------------------------------------
def make_graph(kind, n, rng):
    if kind == "sym_binary_big":       # large enough that every operator of the case runs the merge-path tiled kernel
        m = sp.random(n, n, density=20.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k))
        a = ((m + m.T) > 0).astype(np.float64)
        a.setdiag(0)
    elif kind == "sym_binary":
        m = sp.random(n, n, density=6.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k))
        a = ((m + m.T) > 0).astype(np.float64)
        a.setdiag(0)
    elif kind == "asym_weighted":
        a = sp.random(n, n, density=5.0 / n, random_state=rng, data_rvs=lambda k: rng.integers(1, 4, k).astype(np.float64))
        a = sp.lil_matrix(a)
        a.setdiag(0)
        a[3, 3] = 2.0  # one explicit self loop in A
    elif kind == "isolated":
        m = sp.random(n, n, density=4.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k))
        a = sp.lil_matrix(((m + m.T) > 0).astype(np.float64))
        a.setdiag(0)
        for r in (0, 7, n - 1):
            a[r, :] = 0
            a[:, r] = 0
    elif kind == "hub":
        m = sp.random(n, n, density=3.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k))
        a = sp.lil_matrix(((m + m.T) > 0).astype(np.float64))
        a.setdiag(0)
        a[5, :] = 1
        a[:, 5] = 1
        a[5, 5] = 0
    else:
        raise ValueError(kind)
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    return a


def full_batch_case(ref_model, name, kind, n, d, h, seed, mean, var):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    random.seed(seed)
    a = make_graph(kind, n, rng)
    x = rng.standard_normal((n, d)).astype(np.float32)
    # run.py:96-109
    adj = (_ref_normalize_adj(a) + sp.eye(n)).todense()
    raw_adj = (a + sp.eye(n)).todense()
    adj_t = torch.FloatTensor(np.asarray(adj)[np.newaxis])
    raw_t = torch.FloatTensor(np.asarray(raw_adj)[np.newaxis])
    feats = torch.FloatTensor(x[np.newaxis])
    model = ref_model.Model(d, h, "prelu", 1, "avg")
    with torch.no_grad():  # non-trivial bias / slope so they are exercised
        for k, v in model.state_dict().items():
            if k.endswith("bias"):
                v.copy_(torch.randn_like(v) * 0.1)
        model.gcn2.act.weight.fill_(0.1)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    perm = rng.permutation(n).tolist()
    normal_idx = perm[: max(8, n // 4)]
    abnormal_idx = normal_idx[: max(3, len(normal_idx) // 5)]
    args = types.SimpleNamespace(mean=mean, var=var)
    # Model.forward draws the noise internally (model.py:143); reproduce the draw
    gen_state = torch.get_rng_state()
    noise = torch.randn(1, len(abnormal_idx), h) * var + mean
    torch.set_rng_state(gen_state)
    emb, emb_combine, logits, emb_con, emb_abnormal = model(feats, adj_t, abnormal_idx, normal_idx, True, args)
    # ---- run.py:164-210 on the dense tensors ----
    b_xent = nn.BCEWithLogitsLoss(reduction="none", pos_weight=torch.tensor([1]))
    lbl = torch.unsqueeze(torch.cat((torch.zeros(len(normal_idx)), torch.ones(len(emb_con)))), 1).unsqueeze(0)
    loss_bce = torch.mean(b_xent(logits, lbl))
    e = torch.squeeze(emb)
    e_inf = torch.pow(torch.norm(e, dim=-1, keepdim=True), -1)
    e_inf[torch.isinf(e_inf)] = 0.0
    e_norm = e * e_inf
    sim = torch.mm(e_norm, e_norm.T)
    raw = torch.squeeze(raw_t)
    r_inv = torch.pow(torch.sum(raw, 0), -1)
    r_inv[torch.isinf(r_inv)] = 0.0
    affinity = torch.sum(sim * raw, 0) * r_inv
    margin = (0.7 - (torch.mean(affinity[normal_idx]) - torch.mean(affinity[abnormal_idx]))).clamp_min(min=0)
    rec = torch.mean(torch.sqrt(torch.sum(torch.pow(emb_con - emb_abnormal, 2), 1)))
    loss = margin + loss_bce + rec
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in model.named_parameters()}
    with torch.no_grad():
        gen_state = torch.get_rng_state()
        ev = model(feats, adj_t, abnormal_idx, normal_idx, False, args)
        torch.set_rng_state(gen_state)
        sp_out = model.gcn1(feats, torch.squeeze(adj_t, 0).to_sparse(), sparse=True)  # model.py:28-29 branch
    a = a.tocsr()
    out = dict(
        kind=kind, n=n, d=d, h=h, seed=seed, mean=mean, var=var,
        a_indptr=a.indptr.astype(np.int64), a_indices=a.indices.astype(np.int32), a_data=a.data.astype(np.float64),
        x=x, normal_idx=np.asarray(normal_idx, np.int64), abnormal_idx=np.asarray(abnormal_idx, np.int64),
        noise=noise.squeeze(0).numpy(),
    )
    for k, v in state.items():
        out["p/" + k] = v.numpy()
    for k, v in grads.items():
        out["g/" + k] = v.numpy()
    out.update({
        "o/emb": emb.detach().squeeze(0).numpy(), "o/emb_combine": emb_combine.detach().squeeze(0).numpy(),
        "o/logits": logits.detach().squeeze(0).numpy(), "o/emb_con": emb_con.detach().numpy(),
        "o/emb_abnormal": emb_abnormal.detach().squeeze(0).numpy(), "o/affinity": affinity.detach().numpy(),
        "o/loss": loss.detach().numpy(), "o/margin": margin.detach().numpy(), "o/bce": loss_bce.detach().numpy(),
        "o/rec": rec.detach().numpy(), "o/eval_emb": ev[0].squeeze(0).numpy(), "o/eval_logits": ev[2].squeeze(0).numpy(),
        "o/gcn1_sparse": sp_out.squeeze(0).numpy(),
        "o/adj_hat_dense": np.asarray(adj, dtype=np.float32), "o/deg_rowsum": np.asarray(a.sum(1)).reshape(-1),
    })
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "loss", float(loss), "margin", float(margin))


# ----------------------------------------------------------------------------
def make_adj_lists(n, avg_deg, rng, hub=None, isolated=()):
    from collections import defaultdict
    adj = defaultdict(set)
    m = int(n * avg_deg / 2)
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    for s, t in zip(src.tolist(), dst.tolist()):
        if s == t or s in isolated or t in isolated:
            continue
        adj[s].add(t)
        adj[t].add(s)
    if hub is not None:
        for t in range(n):
            if t != hub and t not in isolated and rng.random() < 0.6:
                adj[hub].add(t)
                adj[t].add(hub)
    for v in range(n):  # every non-isolated node has >= 1 neighbor so adj_list.get() never returns None
        if v in isolated:
            continue
        if len(adj[v]) == 0:
            t = (v + 1) % n
            while t in isolated:
                t = (t + 1) % n
            adj[v].add(t)
            adj[t].add(v)
    return adj


def minibatch_case(ref_sage, name, n, d, h, bsz, n_ab, seed, hub=None, dup=False, avg_deg=5.0):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    random.seed(seed)
    adj = make_adj_lists(n, avg_deg, rng, hub=hub)
    x = rng.random((n, d)).astype(np.float32)
    feats = nn.Embedding(n, d)
    feats.weight = nn.Parameter(torch.FloatTensor(x), requires_grad=False)   # model_handler.py:263-264
    agg = ref_sage.GCNAggregator(feats, cuda=False)
    enc = ref_sage.GCNEncoder(feats, d, h, adj, agg, gcn=True, cuda=False)
    model = ref_sage.GCN(2, enc)
    nodes = rng.permutation(n)[:bsz].tolist()
    if dup:
        nodes[-1] = nodes[0]
    labels = np.zeros(bsz, dtype=np.int64)
    labels[rng.permutation(bsz)[:n_ab]] = 1
    lab_t = torch.LongTensor(labels)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    total, cls, margin, rec = model.loss(nodes, lab_t)
    total.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
    with torch.no_grad():
        prob = model.to_prob(nodes, None)
        to_feats, to_feats_neigh, mask_row = agg.forward(nodes, [adj[int(v)] for v in nodes], adj, True)
        # recover the reference's frontier order to make rows comparable
        samp = [adj[int(v)].union({int(v)}) for v in nodes]
        u_list = list(set.union(*samp))
        embeds, ego, af, afn = enc(nodes, lab_t, True)
    keys = sorted(adj.keys())
    rowptr = np.zeros(n + 1, dtype=np.int64)
    for k in keys:
        rowptr[k + 1] = len(adj[k])
    np.cumsum(rowptr, out=rowptr)
    col = np.concatenate([np.array(sorted(adj[k]), dtype=np.int32) for k in range(n)])
    out = dict(n=n, d=d, h=h, seed=seed, adj_rowptr=rowptr, adj_col=col, x=x,
               nodes=np.asarray(nodes, np.int64), labels=labels)
    for k, v in state.items():
        if k.startswith("enc.features"):
            continue
        out["p/" + k] = v.numpy()
    for k, v in grads.items():
        out["g/" + k] = v.numpy()
    out.update({
        "o/total": total.detach().numpy(), "o/cls": cls.detach().numpy(), "o/margin": margin.detach().numpy(),
        "o/rec": rec.detach().numpy(), "o/prob": prob.numpy(), "o/to_feats": to_feats.numpy(),
        "o/to_feats_neigh": to_feats_neigh.numpy(), "o/mask_row": mask_row.numpy(),
        "o/u_list": np.asarray(u_list, np.int64), "o/embeds": embeds.numpy(), "o/ego": ego.numpy(),
        "o/anomaly_feat": af.numpy(), "o/anomaly_feat_new": afn.numpy(),
    })
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "total", float(total))


def sage_case(ref_sage, name, n, d, h, bsz, seed, gcn):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    adj = make_adj_lists(n, 4.0, rng)
    x = rng.random((n, d)).astype(np.float32)
    feats = nn.Embedding(n, d)
    feats.weight = nn.Parameter(torch.FloatTensor(x), requires_grad=False)
    agg = ref_sage.MeanAggregator(feats, cuda=False, gcn=gcn)
    enc = ref_sage.Encoder(feats, d, h, adj, agg, num_sample=None, gcn=gcn, cuda=False)
    model = ref_sage.GraphSage(2, enc)
    nodes = rng.permutation(n)[:bsz].tolist()
    labels = torch.LongTensor(rng.integers(0, 2, bsz))
    loss = model.loss(nodes, labels)
    loss.backward()
    with torch.no_grad():
        mean = agg.forward(nodes, [adj[int(v)] for v in nodes], None)
        emb = enc(nodes)
        scores = model(nodes)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    for k in range(n):
        rowptr[k + 1] = len(adj[k])
    np.cumsum(rowptr, out=rowptr)
    col = np.concatenate([np.array(sorted(adj[k]), dtype=np.int32) for k in range(n)])
    out = dict(n=n, d=d, h=h, seed=seed, gcn=gcn, adj_rowptr=rowptr, adj_col=col, x=x,
               nodes=np.asarray(nodes, np.int64), labels=labels.numpy())
    out["p/enc.weight"] = enc.weight.detach().numpy()
    out["p/weight"] = model.weight.detach().numpy()
    out["g/enc.weight"] = enc.weight.grad.numpy()
    out["g/weight"] = model.weight.grad.numpy()
    out.update({"o/mean": mean.numpy(), "o/emb": emb.numpy(), "o/scores": scores.numpy(), "o/loss": loss.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "loss", float(loss))


def _load_ref_utils():
    """utils.py itself, with empty stand-ins for the plotting / dgl imports it does not need for load_mat."""
    for name in ("dgl", "matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "seaborn"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["matplotlib.backends.backend_pdf"].__dict__.setdefault("PdfPages", object)
    sys.modules["matplotlib"].__dict__.setdefault("use", lambda *a, **k: None)          # utils.py:179-181 run at import
    sys.modules["matplotlib.pyplot"].__dict__.setdefault("rcParams", {})
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)
    return ref_utils


def data_case(ref_utils, name, dataset, n, d, seed, alt_keys, with_kinds):
    """A synthetic .mat in the published layout -> the reference's load_mat / preprocess_features / normalize_adj."""
    import contextlib
    import io
    import tempfile

    import scipy.io as sio
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, density=5.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k))
    network = sp.csr_matrix(((m + m.T) > 0).astype(np.float64))
    attrs = sp.random(n, d, density=0.3, random_state=rng, data_rvs=lambda k: rng.integers(1, 5, k).astype(np.float64)).tolil()
    attrs[5, :] = 0                                       # a node without attributes: 1 / rowsum = inf -> 0
    label = (rng.random(n) < 0.08).astype(np.int64).reshape(n, 1)
    mat = {("gnd" if alt_keys else "Label"): label, ("X" if alt_keys else "Attributes"): sp.csc_matrix(attrs),
           ("A" if alt_keys else "Network"): sp.csc_matrix(network)}
    if with_kinds:
        str_l = (label[:, 0] * (rng.random(n) < 0.5)).astype(np.int64).reshape(1, n)
        mat["str_anomaly_label"] = str_l
        mat["attr_anomaly_label"] = (label[:, 0].reshape(1, n) - str_l)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "dataset"))
        sio.savemat(os.path.join(tmp, "dataset", dataset + ".mat"), mat)
        os.chdir(tmp)
        try:
            random.seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                out = ref_utils.load_mat(dataset)
        finally:
            os.chdir(cwd)
    adj, feat, ano, all_idx, tr, va, te, ano2, str_a, attr_a, normal, abnormal = out
    dense_feat, (coords, values, shape) = ref_utils.preprocess_features(feat)
    store = {"i/network": network.toarray(), "i/attrs": attrs.toarray(), "i/label": label, "i/seed": seed,
             "i/alt_keys": int(alt_keys), "i/dataset": dataset,
             "o/adj": adj.toarray(), "o/feat": feat.toarray(), "o/ano_labels": ano, "o/all_idx": all_idx, "o/idx_train": tr,
             "o/idx_val": va, "o/idx_test": te, "o/normal_label_idx": normal, "o/abnormal_label_idx": abnormal,
             "o/pre_dense": np.asarray(dense_feat), "o/pre_coords": coords, "o/pre_values": values, "o/pre_shape": np.asarray(shape),
             "o/normalize_adj": ref_utils.normalize_adj(adj).toarray()}
    if with_kinds:
        store.update({"i/str": mat["str_anomaly_label"], "i/attr": mat["attr_anomaly_label"], "o/str": str_a, "o/attr": attr_a})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
    print(f"{name}: n={n} train={len(tr)} normal={len(normal)} abnormal={len(abnormal)}")


def minibatch_data_case(name, n, d, seed):
    """src/utils.py (stub for dgl): normalize, sparse_to_adjlist (the later definition, src/utils.py:96-112, wins at
    import), pos_neg_split on a synthetic graph."""
    import pickle
    import tempfile
    if "dgl" not in sys.modules:
        try:
            __import__("dgl")
        except Exception:
            sys.modules["dgl"] = types.ModuleType("dgl")
    spec = importlib.util.spec_from_file_location("ref_src_utils", os.path.join(REF, "src", "utils.py"))
    u = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(u)
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=4.0 / n, random_state=rng, data_rvs=lambda k: np.ones(k)).tocsr()   # directed, with isolated nodes
    x = rng.random((n, d))
    x[3] = 0
    x[4] = -0.01 / d                                       # row sum exactly -0.01: 1 / 0 = inf -> 0
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "adj")
        u.sparse_to_adjlist(a, fn)
        with open(fn, "rb") as f:
            adj_lists = pickle.load(f)
    keys = np.array(sorted(adj_lists), dtype=np.int64)
    lens = np.array([len(adj_lists[k]) for k in keys], dtype=np.int64)
    flat = np.concatenate([np.array(sorted(adj_lists[k]), dtype=np.int64) for k in keys])
    nodes = rng.permutation(n)[: n // 2].tolist()
    nodes[5] = nodes[2]                                    # a duplicated id
    labels = (rng.random(len(nodes)) < 0.3).astype(np.int64)
    pos, neg = u.pos_neg_split(nodes, labels)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{
        "i/a_data": a.data, "i/a_indices": a.indices, "i/a_indptr": a.indptr, "i/n": n, "i/x": x, "i/nodes": nodes, "i/labels": labels,
        "o/normalize_dense": np.asarray(u.normalize(x)), "o/normalize_sparse": u.normalize(sp.csr_matrix(x)).toarray(),
        "o/adj_keys": keys, "o/adj_lens": lens, "o/adj_flat": flat, "o/pos": pos, "o/neg": neg})
    print(f"{name}: n={n} adjacency-list keys={len(keys)} entries={len(flat)} pos={len(pos)} neg={len(neg)}")


def tam_case(name, kind, n, h, seed):
    """tam.py is a script; its two affinity functions (tam.py:113-146) are taken out of its source with ``ast`` and run
    unchanged on a dense adjacency (R = A + I as tam.py feeds them)."""
    import ast
    tree = ast.parse(open(os.path.join(REF, "tam.py")).read())
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("max_message", "inference"):
            exec(compile(ast.Module([node], []), "tam.py", "exec"), ns)
    rng = np.random.default_rng(seed)
    a = sp.csr_matrix(make_graph(kind, n, rng))
    r = (a + sp.eye(n)).tocsr()
    torch.manual_seed(seed)
    feat = torch.randn(n, h)
    normal = sorted(rng.permutation(n)[: n // 3].tolist())
    dense = torch.from_numpy(r.toarray().astype(np.float32))
    f1 = feat.clone().requires_grad_(True)
    loss, msg = ns["max_message"](f1, dense, normal)
    loss.backward()
    inf = ns["inference"](feat, dense)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{
        "n": n, "r_data": r.data, "r_indices": r.indices, "r_indptr": r.indptr, "feat": feat.numpy(), "normal": normal,
        "o/loss": loss.detach().numpy(), "o/message": msg.detach().numpy(), "o/inference": inf.numpy(), "g/feat": f1.grad.numpy()})
    print(f"{name}: n={n} nnz={r.nnz} loss={float(loss):.5f}")


def encoder_case(name, kind, n, d, h, seed):
    """model_ocgnn.py (imports only torch): its two-layer GCN encoder ``Model`` on the dense A_hat run.py builds."""
    spec = importlib.util.spec_from_file_location("ref_model_ocgnn", os.path.join(REF, "model_ocgnn.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(seed)
    a = sp.csr_matrix(make_graph(kind, n, rng))
    a_hat = (_ref_normalize_adj(a) + sp.eye(n)).todense()
    torch.manual_seed(seed)
    m = ref.Model(d, h, "prelu", 1, "avg")
    x = torch.randn(1, n, d)
    out = m(x, torch.FloatTensor(a_hat[np.newaxis]))
    (out * out).sum().mul(0.5).backward()
    store = {"n": n, "d": d, "h": h, "a_data": a.data, "a_indices": a.indices, "a_indptr": a.indptr, "x": x.numpy(), "o/h2": out.detach().numpy()}
    for k, v in m.state_dict().items():
        store["p/" + k] = v.numpy()
    for k, p in m.named_parameters():
        store["g/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
    print(f"{name}: n={n} |h2|={float(out.norm()):.4f}")


def main():
    ref_model, ref_sage = _load_ref()
    encoder_case("enc_sym", "sym_binary", 150, 20, 32, 0)
    encoder_case("enc_asym_weighted", "asym_weighted", 90, 12, 16, 72)
    tam_case("tam_sym", "sym_binary", 120, 16, 0)
    tam_case("tam_asym_weighted", "asym_weighted", 90, 12, 72)
    ref_utils = _load_ref_utils()
    minibatch_data_case("mbdata_toy", 500, 9, 4)
    data_case(ref_utils, "data_toy", "toy", 400, 12, 3, alt_keys=False, with_kinds=True)
    data_case(ref_utils, "data_amazon_keys", "Amazon", 900, 9, 11, alt_keys=True, with_kinds=False)
    full_batch_case(ref_model, "fb_sym_binary", "sym_binary", 64, 12, 16, 0, 0.02, 0.01)
    full_batch_case(ref_model, "fb_asym_weighted", "asym_weighted", 48, 10, 16, 72, 0.0, 0.0)
    full_batch_case(ref_model, "fb_isolated", "isolated", 56, 25, 20, 0, 0.02, 0.01)
    full_batch_case(ref_model, "fb_hub", "hub", 96, 17, 32, 72, 0.0, 0.0)
    full_batch_case(ref_model, "fb_wide", "sym_binary", 80, 48, 128, 0, 0.02, 0.01)
    minibatch_case(ref_sage, "mb_basic", 300, 17, 64, 40, 10, 72)
    minibatch_case(ref_sage, "mb_hub", 240, 10, 16, 32, 8, 0, hub=11)
    minibatch_case(ref_sage, "mb_dup", 200, 25, 32, 24, 6, 72, dup=True)
    # cases whose operators cross ggad_b200.graph.PLAN_MIN_ITEMS (rows + edges >= 16 384), so the reference-derived
    # goldens reach gather_tiled_kernel and not only the row-per-group kernel
    full_batch_case(ref_model, "fb_big", "sym_binary_big", 3000, 24, 32, 0, 0.02, 0.01)
    minibatch_case(ref_sage, "mb_big", 6000, 17, 64, 600, 150, 72, avg_deg=10.0)
    sage_case(ref_sage, "sage_concat", 150, 17, 32, 20, 0, gcn=False)
    sage_case(ref_sage, "sage_gcn", 150, 10, 16, 20, 72, gcn=True)


if __name__ == "__main__":
    main()
