"""Host-side integer logic (frontier blocks, degrees, partitioner, preprocessing) -- CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from helpers import case_adj_lists, case_adjacency, golden_cases, load_case


@pytest.mark.parametrize("name", golden_cases("mb_"))
def test_batch_blocks_bit_exact_vs_oracle_and_reference(name):
    from ggad_b200.graph import AdjListCSR, batch_block
    c = load_case(name)
    adj = case_adj_lists(c)
    nodes = c["nodes"].tolist()
    acsr = AdjListCSR(adj)
    hop1 = batch_block(acsr, nodes, add_self=True)
    ref = oracle.gcn_aggregator(nodes, adj, torch.from_numpy(c["x"]), True)
    assert hop1["frontier"].tolist() == ref["U"] == sorted(c["out"]["u_list"].tolist())
    assert np.array_equal(hop1["rdeg"], ref["rdeg"]) and np.array_equal(hop1["cdeg"], ref["cdeg"])
    rows = np.repeat(np.arange(len(nodes)), hop1["rdeg"])
    assert np.array_equal(rows, ref["rows"]) and np.array_equal(hop1["col"], ref["cols"])
    hop2 = batch_block(acsr, hop1["frontier"], add_self=False)
    assert hop2["frontier"].tolist() == ref["U2"]
    assert np.array_equal(hop2["rdeg"], ref["rdeg2"]) and np.array_equal(hop2["cdeg"], ref["cdeg2"])
    # and against the reference's own dense masks
    order = np.argsort(c["out"]["u_list"])
    mask = c["out"]["mask_row"][:, order] > 0
    assert np.array_equal(hop1["rdeg"], mask.sum(1)) and np.array_equal(hop1["cdeg"], mask.sum(0))


def test_adjlist_csr_edge_cases():
    from ggad_b200.graph import AdjListCSR, batch_block
    adj = {0: {1, 2}, 1: {0}, 2: {0, 3}, 3: {2}, 5: set()}
    a = AdjListCSR(adj)
    b = batch_block(a, [0, 3, 5, 0, 9], add_self=True)         # duplicate batch node, isolated node, unknown id
    assert b["frontier"].tolist() == [0, 1, 2, 3, 5, 9]
    assert b["rdeg"].tolist() == [3, 2, 1, 3, 1]
    assert b["cdeg"].tolist() == [2, 2, 3, 1, 1, 1]
    b2 = batch_block(a, [5, 1], add_self=False)
    assert b2["rdeg"].tolist() == [0, 1] and b2["frontier"].tolist() == [0]
    assert AdjListCSR.get(adj) is AdjListCSR.get(adj)


@pytest.mark.parametrize("name", golden_cases("fb_"))
def test_normalize_adj_bit_exact(name):
    from ggad_b200.graph import normalize_adj_scipy
    c = load_case(name)
    a = case_adjacency(c)
    n = a.shape[0]
    a_hat = (normalize_adj_scipy(a) + sp.eye(n)).astype(np.float32)
    assert np.array_equal(np.asarray(a_hat.todense()), c["out"]["adj_hat_dense"])
    ref_hat, _ = oracle.build_full_batch_graph(a)
    assert np.array_equal(np.asarray(a_hat.todense()), np.asarray(ref_hat.todense()))


def test_nnz_balanced_ranges():
    from ggad_b200.dist import even_ranges, nnz_balanced_ranges
    rng = np.random.default_rng(0)
    deg = np.minimum((rng.pareto(1.2, 10000) * 3).astype(np.int64), 4000)
    rowptr = np.zeros(10001, np.int64)
    np.cumsum(deg, out=rowptr[1:])
    for world in (1, 2, 4, 8):
        r = nnz_balanced_ranges(rowptr, world)
        assert r[0][0] == 0 and r[-1][1] == 10000
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        nnz = [rowptr[hi] - rowptr[lo] for lo, hi in r]
        assert sum(nnz) == rowptr[-1]
        assert max(nnz) - min(nnz) <= 2 * deg.max()
        ri = nnz_balanced_ranges(rowptr, world, row_cost=1)                    # merge-path items (rows + edges)
        assert ri[0][0] == 0 and ri[-1][1] == 10000 and all(ri[i][1] == ri[i + 1][0] for i in range(world - 1))
        items = [rowptr[hi] - rowptr[lo] + hi - lo for lo, hi in ri]
        assert sum(items) == rowptr[-1] + 10000 and max(items) - min(items) <= 2 * (deg.max() + 1)
    assert even_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert nnz_balanced_ranges(np.zeros(6, np.int64), 3)[-1][1] == 5          # empty graph still partitions


def test_fit_row_cost_recovers_the_cost_model():
    """Profile-guided split: per-rank (rows, nnz, seconds) samples -> cost of a row in edges."""
    from ggad_b200.dist import fit_row_cost, nnz_balanced_ranges
    a, b = 2.0e-8, 5.0e-8
    shards = [(300_000, 142_000_000), (5_800_000, 126_000_000), (18_700_000, 87_000_000)]
    samples = [(r, m, a * m + b * r) for r, m in shards]
    assert abs(fit_row_cost(samples) - b / a) < 1e-6
    assert fit_row_cost(samples[:1]) is None                               # one sample determines nothing
    assert fit_row_cost([(10, 100, 1.0), (20, 200, 2.0)]) is None           # collinear shapes
    assert fit_row_cost([(r, m, 1e-8 * m + 1e-6 * r) for r, m in shards]) == 8.0     # clipped
    # fractional costs stay exact-integer splits and identical for identical inputs
    rowptr = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(0, 50, 1000))]).astype(np.int64)
    r1 = nnz_balanced_ranges(rowptr, 4, row_cost=1.8125)
    assert r1 == nnz_balanced_ranges(rowptr.copy(), 4, row_cost=1.8125) and r1[-1][1] == 1000


def test_reshard_rows_index_arithmetic():
    """dist.reshard_rows: compute-balanced source ranges -> owners' even ranges, every row lands exactly once
    (the PeerBlock is replaced by local tensors; only the overlap arithmetic is under test)."""
    from ggad_b200.dist import even_ranges, nnz_balanced_ranges, reshard_rows
    n, d, world = 1000, 3, 4
    rng = np.random.default_rng(3)
    rowptr = np.concatenate([[0], np.cumsum(rng.pareto(1.1, n).astype(np.int64) + 1)])
    src = nnz_balanced_ranges(rowptr, world, row_cost=1.5)          # very uneven row counts
    dst = even_ranges(n, world)
    full = torch.arange(n * d, dtype=torch.float32).reshape(n, d)

    class Block:
        def __init__(self):
            self.bufs = [torch.full((hi - lo, d), float("nan")) for lo, hi in dst]

        def view(self, p):
            return self.bufs[p]
    blk = Block()
    for r in range(world):
        lo, hi = src[r]
        reshard_rows(full[lo:hi], (lo, hi), dst, blk)
    for p, (lo, hi) in enumerate(dst):
        assert torch.equal(blk.bufs[p], full[lo:hi])
    assert len({hi - lo for lo, hi in src}) > 1


def test_halo_need_mask_single_rank_is_empty():
    from ggad_b200.dist import halo_need_mask
    m = halo_need_mask(torch.tensor([0, 3, 3, 7], dtype=torch.int32), [(0, 10)], 0)
    assert m.dtype == torch.int32 and m.shape == (10,) and int(m.abs().sum()) == 0


def test_planted_graph_and_power_law_generators():
    from ggad_b200 import synth
    a, x, y = synth.planted_anomaly_graph(800, 10.0, 16, 0.08, seed=1)
    assert (abs(a - a.T)).nnz == 0 and a.diagonal().sum() == 0 and x.shape == (800, 16)
    assert 20 < y.sum() < 120
    adj = synth.power_law_adj_lists(500, 6.0, seed=2)
    assert all(v in adj[u] for u in adj for v in adj[u]) and all(u not in adj[u] for u in adj)
    assert max(len(s) for s in adj.values()) > 8 * np.mean([len(s) for s in adj.values()])


def test_device_metrics_match_sklearn_on_cpu_tensors():
    """metrics.py is plain tensor code (sort + prefix sums): checked here on CPU tensors against sklearn,
    including heavy ties; the GPU suite re-runs it on CUDA tensors."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from ggad_b200 import metrics
    rng = np.random.default_rng(0)
    for n, ties in ((500, False), (2000, True), (50, True)):
        y = (rng.random(n) < 0.15).astype(np.int64)
        y[:2] = [0, 1]
        s = rng.standard_normal(n) + y * 0.8
        if ties:
            s = np.round(s, 1)
        auc = float(metrics.roc_auc(torch.from_numpy(s), torch.from_numpy(y)))
        ap = float(metrics.average_precision(torch.from_numpy(s), torch.from_numpy(y)))
        assert abs(auc - roc_auc_score(y, s)) < 1e-12
        assert abs(ap - average_precision_score(y, s)) < 1e-12
