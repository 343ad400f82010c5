"""Pin the CPU oracle against vectors produced by the reference modules themselves
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, case_adj_lists, case_adjacency, golden_cases, load_case


@pytest.mark.parametrize("name", golden_cases("fb_"))
def test_full_batch_oracle_matches_reference(name):
    c = load_case(name)
    a = case_adjacency(c)
    a_hat, r = oracle.build_full_batch_graph(a)
    # index / degree work: bit-exact
    assert np.array_equal(np.asarray(a.sum(1)).reshape(-1), c["out"]["deg_rowsum"])
    dense = np.asarray(a_hat.todense(), dtype=np.float32)
    assert np.array_equal(dense, c["out"]["adj_hat_dense"]), "A_hat must be bit-exact (fp64 product, one fp32 rounding)"
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["params"].items()}
    x = torch.from_numpy(c["x"])
    res = oracle.full_batch_step(p, x, oracle.csr_arrays(a_hat), oracle.csr_arrays(r),
                                 c["abnormal_idx"].tolist(), c["normal_idx"].tolist(), torch.from_numpy(c["noise"]))
    o = c["out"]
    for k in ("emb", "emb_combine", "logits", "emb_con", "emb_abnormal"):
        assert_close(res[k], o[k], what=f"{name}:{k}")
    sub = np.concatenate([c["normal_idx"], c["abnormal_idx"]])
    assert_close(res["affinity"][sub], o["affinity"][sub], what="affinity")
    assert_close(res["affinity"], o["affinity"], what="affinity(all)")
    for k in ("loss", "margin", "bce", "rec"):
        assert_close(res[k], o[k], what=k)
    res["loss"].backward()
    for k, g in c["grads"].items():
        got = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        assert_close(got, g, rtol=2e-4, atol=1e-6, what=f"{name}:grad {k}")
    ev = oracle.model_forward({k: v.detach() for k, v in p.items()}, x, oracle.csr_arrays(a_hat),
                              c["abnormal_idx"].tolist(), c["normal_idx"].tolist(), False, torch.from_numpy(c["noise"]))
    assert_close(ev[0], o["eval_emb"], what="eval emb")
    assert_close(ev[2], o["eval_logits"], what="eval logits")
    g1 = oracle.gcn_layer(x, oracle.csr_arrays(a_hat), p["gcn1.fc.weight"].detach(), p["gcn1.bias"].detach(),
                          p["gcn1.act.weight"].detach())
    assert_close(g1, o["gcn1_sparse"], what="model.py:28-29 sparse branch")


@pytest.mark.parametrize("name", golden_cases("mb_"))
def test_minibatch_oracle_matches_reference(name):
    c = load_case(name)
    adj = case_adj_lists(c)
    feats = torch.from_numpy(c["x"])
    nodes = c["nodes"].tolist()
    labels = torch.from_numpy(c["labels"])
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["params"].items()}
    o = c["out"]
    agg = oracle.gcn_aggregator(nodes, adj, feats, True)
    # frontier: same SET as the reference, integer degrees exact
    assert sorted(o["u_list"].tolist()) == agg["U"]
    ref_mask = o["mask_row"]
    order = np.argsort(o["u_list"])                       # reference (set-order) columns -> sorted
    assert np.array_equal((ref_mask[:, order] > 0).sum(1), agg["rdeg"])
    assert np.array_equal((ref_mask[:, order] > 0).sum(0), agg["cdeg"])
    dense = np.zeros_like(ref_mask)
    dense[agg["rows"], agg["cols"]] = agg["mask_row_w"].numpy()
    assert np.array_equal(dense, ref_mask[:, order]), "mean mask must be bit-exact"
    assert_close(agg["to_feats"], o["to_feats"], what="to_feats")
    assert_close(agg["to_feats_neigh"], o["to_feats_neigh"][order], what="to_feats_neigh")
    total, cls, margin, rec = oracle.gcn_minibatch_loss(p, nodes, labels, adj, feats)
    for k, v in (("total", total), ("cls", cls), ("margin", margin), ("rec", rec)):
        assert_close(v, o[k], what=k)
    total.backward()
    for k, g in c["grads"].items():
        assert_close(p[k].grad, g, rtol=2e-4, atol=1e-6, what=f"grad {k}")
    pd = {k: v.detach() for k, v in p.items()}
    scores = oracle.gcn_minibatch_forward(pd, nodes, None, adj, feats, False)[0]
    assert_close(torch.sigmoid(scores), o["prob"], what="to_prob")
    emb, ego, af, afn = oracle.gcn_encoder(pd, nodes, labels, adj, feats, True)
    assert_close(emb, o["embeds"], what="combined_all")
    assert_close(ego, o["ego"], what="ego")
    assert_close(af, o["anomaly_feat"], what="anomaly_feat")
    assert_close(afn, o["anomaly_feat_new"], what="anomaly_feat_new")


@pytest.mark.parametrize("name", golden_cases("sage_"))
def test_sage_oracle_matches_reference(name):
    c = load_case(name)
    adj = case_adj_lists(c)
    feats = torch.from_numpy(c["x"])
    nodes = c["nodes"].tolist()
    gcn = bool(c["gcn"])
    mean = oracle.mean_aggregator(nodes, [adj[n] for n in nodes], feats, gcn=gcn)
    assert_close(mean, c["out"]["mean"], what="mean")
    w = c["params"]["enc.weight"].clone().requires_grad_(True)
    cw = c["params"]["weight"].clone().requires_grad_(True)
    emb = oracle.sage_encoder(w, nodes, adj, feats, gcn=gcn)
    assert_close(emb, c["out"]["emb"], what="emb")
    scores = cw.mm(emb).t()
    assert_close(scores, c["out"]["scores"], what="scores")
    loss = torch.nn.functional.cross_entropy(scores, torch.from_numpy(c["labels"]))
    assert_close(loss, c["out"]["loss"], what="loss")
    loss.backward()
    assert_close(w.grad, c["grads"]["enc.weight"], rtol=2e-4, atol=1e-6, what="grad enc.weight")
    assert_close(cw.grad, c["grads"]["weight"], rtol=2e-4, atol=1e-6, what="grad weight")


def test_isolated_nodes_nan_like_reference():
    adj = {0: {1}, 1: {0}, 2: set()}
    feats = torch.arange(9, dtype=torch.float32).reshape(3, 3)
    out = oracle.mean_aggregator([0, 2], [adj[0], adj[2]], feats)
    assert torch.isnan(out[1]).all() and not torch.isnan(out[0]).any()
    agg = oracle.gcn_aggregator([2, 0], adj, feats, True)
    u = agg["U"]
    assert torch.isnan(agg["to_feats_neigh"][u.index(2)]).all()


def test_preprocessing_restated():
    rng = np.random.default_rng(0)
    x = rng.random((5, 4))
    x[2] = 0
    out = oracle.preprocess_features(x)
    assert np.allclose(out[0].sum(), 1.0) and np.all(out[2] == 0)
    mb = oracle.normalize_rows_minibatch(x)
    assert np.allclose(mb[0], x[0] / (x[0].sum() + 0.01))
    labels = (rng.random(200) < 0.1).astype(int)
    tr, va, te, normal, abnormal = oracle.load_mat_split(labels, "photo", 0)
    assert len(tr) == 60 and len(va) == 20 and len(te) == 120
    assert set(abnormal) <= set(normal) and all(labels[i] == 0 for i in normal)
    assert len(abnormal) == int(len(normal) * 0.15)
