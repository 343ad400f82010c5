"""Property tests (hypothesis) that pin the CSR oracle to the reference's DENSE formulations, written out here the way
run.py / model.py / src/graphsage.py write them (dense N x N matrices, dense 0/1 masks), over random small graphs:
asymmetric and weighted adjacency, isolated nodes, hubs, duplicate batch ids, widths that are not multiples of 4.
Integer outputs (frontiers, degrees) must be equal; fp32 outputs within 1e-5 (summation order only)."""
import numpy as np
import scipy.sparse as sp
import torch
from hypothesis import given, settings, strategies as st

import oracle

SET = settings(max_examples=40, deadline=None)


@st.composite
def graphs(draw, max_n=40):
    n = draw(st.integers(3, max_n))
    density = draw(st.floats(0.02, 0.4))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    a = (rng.random((n, n)) < density).astype(np.float64)
    if draw(st.booleans()):
        a = np.maximum(a, a.T)                                   # symmetric
    if draw(st.booleans()):
        a *= rng.random((n, n)) * 2.0 + 0.1                      # weighted
    if draw(st.booleans()):
        a[rng.integers(0, n)] = 0.0                              # an isolated source row
    if draw(st.booleans()):
        a[:, rng.integers(0, n)] = 1.0                           # a hub column
    np.fill_diagonal(a, 0.0)
    return a, rng


@SET
@given(graphs())
def test_normalize_adj_equals_dense_formula(g):
    a, _ = g
    got = oracle.normalize_adj(sp.csr_matrix(a)).toarray()
    rowsum = a.sum(1)
    with np.errstate(divide="ignore"):
        dis = np.power(rowsum, -0.5)
    dis[np.isinf(dis)] = 0.0
    want = (a @ np.diag(dis)).T @ np.diag(dis)                   # utils.py:50-54: (A D)^T D in fp64
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


@SET
@given(graphs(), st.integers(1, 9))
def test_spmm_equals_dense_matmul(g, d):
    a, rng = g
    m = sp.csr_matrix(a.astype(np.float32))
    x = torch.from_numpy(rng.standard_normal((a.shape[1], d)).astype(np.float32))
    got = oracle.spmm_csr(m.indptr, m.indices, m.data, x, n_rows=a.shape[0])
    want = torch.from_numpy(a.astype(np.float32)) @ x             # model.py:31 torch.bmm(adj, seq_fts)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


@SET
@given(graphs(), st.integers(1, 9))
def test_local_affinity_equals_dense_formula(g, h):
    a, rng = g
    n = a.shape[0]
    r = a.astype(np.float32) + np.eye(n, dtype=np.float32)      # raw_adj = A + I (run.py:100)
    emb = torch.from_numpy(rng.standard_normal((n, h)).astype(np.float32))
    if n > 4:
        emb[2] = 0.0                                             # zero embedding: 1/0 -> 0 (run.py:178-179)
    rc = sp.csr_matrix(r)
    got = oracle.local_affinity(emb, (rc.indptr, rc.indices, rc.data))
    # run.py:175-188 verbatim on dense tensors
    emb_inf = torch.norm(emb, dim=-1, keepdim=True)
    emb_inf = torch.pow(emb_inf, -1)
    emb_inf[torch.isinf(emb_inf)] = 0.
    emb_norm = emb * emb_inf
    sim = torch.mm(emb_norm, emb_norm.T) * torch.from_numpy(r)
    row_sum = torch.sum(torch.from_numpy(r), 0)
    r_inv = torch.pow(row_sum, -1)
    r_inv[torch.isinf(r_inv)] = 0.
    want = torch.sum(sim, 0) * r_inv
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@SET
@given(graphs(max_n=30), st.integers(1, 8), st.integers(1, 7))
def test_gcn_aggregator_equals_dense_mask_formulation(g, batch, d):
    a, rng = g
    n = a.shape[0]
    und = (a + a.T) > 0
    adj_lists = {i: set(np.flatnonzero(und[i]).tolist()) for i in range(n)}      # no self loops (src/utils.py:96-112)
    nodes = rng.integers(0, n, batch).tolist()                                  # duplicates allowed
    feats = torch.from_numpy(rng.random((n, d)).astype(np.float32))
    out = oracle.gcn_aggregator(nodes, adj_lists, feats, train_flag=True)
    # src/graphsage.py:305-326 on a dense mask (frontier sorted instead of Python-set order)
    samp = [set(adj_lists[v]) | {v} for v in nodes]
    u = sorted(set.union(*samp))
    pos = {v: i for i, v in enumerate(u)}
    mask = torch.zeros(len(samp), len(u))
    for i, s in enumerate(samp):
        for v in s:
            mask[i, pos[v]] = 1
    rdeg, cdeg = mask.sum(1, keepdim=True), mask.sum(0, keepdim=True)
    want = mask.div(rdeg.sqrt()).div(cdeg.sqrt()).mm(feats[torch.tensor(u)])
    assert out["U"] == u or list(out["U"]) == u
    assert np.array_equal(np.asarray(out["rdeg"]), rdeg.flatten().numpy().astype(np.int64))
    assert np.array_equal(np.asarray(out["cdeg"]), cdeg.flatten().numpy().astype(np.int64))
    assert torch.allclose(out["to_feats"], want, rtol=1e-5, atol=1e-6)
    # hop 2 (:335-355): no self union; empty rows are 0/0 = NaN in the dense formulation
    samp2 = [set(adj_lists[v]) for v in u]
    u2 = sorted(set.union(*samp2)) if any(samp2) else []
    if u2:
        pos2 = {v: i for i, v in enumerate(u2)}
        m2 = torch.zeros(len(u), len(u2))
        for i, s in enumerate(samp2):
            for v in s:
                m2[i, pos2[v]] = 1
        m2 = m2.div(m2.sum(1, keepdim=True).sqrt()).div(m2.sum(0, keepdim=True).sqrt())
        want2 = m2.mm(feats[torch.tensor(u2)])
        got2 = out["to_feats_neigh"]
        rows_ok = torch.tensor([len(s) > 0 for s in samp2])
        assert torch.allclose(got2[rows_ok], want2[rows_ok], rtol=1e-5, atol=1e-6)
        assert bool(torch.isnan(got2[~rows_ok]).all())
