"""world_size-2 gloo test of the node-range sharded layer pass (partition + collective logic).
The local compute is injected (the CPU oracle stands in for the CUDA kernel in this test only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import scipy.sparse as sp
    import oracle
    from ggad_b200.dist import ShardedLayerPass, halo_need_mask, nnz_balanced_ranges
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n, d = 300, 12
    deg = np.minimum((rng.pareto(1.2, n) * 3).astype(np.int64), 200)
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=rowptr[1:])
    col = rng.integers(0, n, rowptr[-1]).astype(np.int32)
    val = rng.random(rowptr[-1]).astype(np.float32)
    a = sp.csr_matrix((val, col, rowptr), shape=(n, n))
    at = a.T.tocsr()
    fr = nnz_balanced_ranges(a.indptr, world)
    br = nnz_balanced_ranges(at.indptr, world)

    def shard(m, lo, hi):
        s = m[lo:hi]
        return (s.indptr.astype(np.int64), s.indices.astype(np.int32), s.data.astype(np.float32))

    def compute(g, x):
        return oracle.spmm_csr(g[0], g[1], g[2], x)
    lp = ShardedLayerPass(shard(a, *fr[rank]), shard(at, *br[rank]), fr, br, rank, compute)
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    y = lp.forward(x)
    dx = lp.backward(y)
    y_ref = oracle.spmm_csr(a.indptr, a.indices, a.data.astype(np.float32), x)
    dx_ref = oracle.spmm_csr(at.indptr, at.indices, at.data.astype(np.float32), y_ref)
    ok = torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6) and torch.allclose(dx, dx_ref, rtol=1e-5, atol=1e-5)
    # every rank ends with the same replicated result
    gathered = [torch.empty_like(dx) for _ in range(world)]
    dist.all_gather(gathered, dx)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    # halo masks (integer, exact): bit s of mask[r] <=> the s-th peer's backward shard has row lo+r as a column
    mask = halo_need_mask(torch.from_numpy(shard(at, *br[rank])[1]), fr, rank)
    lo, hi = fr[rank]
    peers = [r for r in range(world) if r != rank]
    expect = np.zeros(hi - lo, np.int32)
    for s, p in enumerate(peers):
        cols = np.unique(shard(at, *br[p])[1])
        cols = cols[(cols >= lo) & (cols < hi)] - lo
        expect[cols] |= 1 << s
    ok = ok and mask.dtype == torch.int32 and np.array_equal(mask.numpy(), expect) and 0 < int((expect != 0).sum())
    q.put((rank, bool(ok), fr, br))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_layer_pass_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]      # identical partition on every rank


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from ggad_b200.train import DataParallelMiniBatch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Tiny(torch.nn.Module):                      # stands in for the drop-in GCN: same .loss() contract
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.w = torch.nn.Parameter(torch.randn(5, 3))
            self.unused = torch.nn.Parameter(torch.randn(2))     # never gets a gradient
            self.frozen = torch.nn.Parameter(torch.randn(2), requires_grad=False)

        def loss(self, nodes, labels):
            x = torch.tensor(nodes, dtype=torch.float32).reshape(-1, 5)
            t = ((x @ self.w).sum(1) - labels.float()).pow(2).mean()
            return t, t, t, t
    batches = [([1., 2, 3, 4, 5, 0, 1, 0, 1, 0], torch.tensor([1, 0])), ([2., 2, 2, 2, 2, 9, 8, 7, 6, 5], torch.tensor([0, 1]))]
    m = Tiny()
    dp = DataParallelMiniBatch(m, torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-2))
    for _ in range(3):
        dp.step(*batches[rank])
    ref = Tiny()
    opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=1e-2)
    for _ in range(3):
        opt.zero_grad()
        (sum(ref.loss(*b)[0] for b in batches) / world).backward()
        if ref.unused.grad is None:
            ref.unused.grad = torch.zeros_like(ref.unused)
        opt.step()
    ok = torch.allclose(m.w, ref.w, rtol=1e-6, atol=1e-7) and torch.equal(m.frozen, ref.frozen)
    q.put((rank, bool(ok), m.w.detach().tolist()))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_data_parallel_minibatch_world2():
    """Gradient averaging of the data-parallel mini-batch driver == one process minimising the mean loss."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]                                   # replicas stay identical


# ------------------------------------------------------------------------------------------
# sharded-feature mini-batch (SURVEY 8e bullet 3): partial accumulators + one all-reduce per layer
# ------------------------------------------------------------------------------------------
class _CpuBackend:
    """torch-CPU stand-ins for the three device primitives of ggad_b200.sharded (this test only)."""

    def __init__(self, rowptr, col):
        self.rowptr, self.col = rowptr, col               # numpy CSR over ALL rows, LOCAL column ids

    def block(self, nodes):
        nodes = nodes.numpy().astype(np.int64)
        lens = self.rowptr[nodes + 1] - self.rowptr[nodes]
        rp = np.zeros(len(nodes) + 1, np.int64)
        np.cumsum(lens, out=rp[1:])
        cols = np.concatenate([self.col[self.rowptr[v]:self.rowptr[v + 1]] for v in nodes]) if len(nodes) else np.zeros(0, np.int32)
        return torch.from_numpy(rp), torch.from_numpy(cols.astype(np.int32))

    def spmm(self, rowptr, col, n_rows, n_cols, table):
        rows = torch.repeat_interleave(torch.arange(n_rows), rowptr[1:] - rowptr[:-1])
        return torch.zeros(n_rows, table.shape[1]).index_add(0, rows, table[col.long()])

    def linear(self, x, w, relu=False):
        y = x @ w.t()
        return torch.relu(y) if relu else y


def _sage_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import oracle
    from ggad_b200.dist import even_ranges
    from ggad_b200.sharded import ShardedTwoLayerSage
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    n, d, h, n_cls = 400, 10, 16, 2
    adj = {v: set() for v in range(n)}                          # simple undirected graph, a hub, no self loops
    for s, t in zip(rng.integers(0, n, 1500).tolist(), rng.integers(0, n, 1500).tolist()):
        if s != t:
            adj[s].add(t); adj[t].add(s)
    for t in range(0, n, 3):
        if t != 7:
            adj[7].add(t); adj[t].add(7)
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    torch.manual_seed(0)
    w = [torch.randn(h, d) * 0.3, torch.randn(h, h) * 0.3, torch.randn(n_cls, h) * 0.3]
    seeds = torch.from_numpy(rng.permutation(n)[:60].astype(np.int64))
    labels = torch.from_numpy(rng.integers(0, n_cls, 60))

    def build(lo, hi):                                          # A[:, lo:hi] as CSR over all rows, local column ids
        rp = np.zeros(n + 1, np.int64)
        cols = []
        for v in range(n):
            c = sorted(t - lo for t in adj[v] if lo <= t < hi)
            cols += c
            rp[v + 1] = rp[v] + len(c)
        return _CpuBackend(rp, np.asarray(cols, np.int32))

    def run(lo, hi, group_world):
        ws = [t.clone().requires_grad_(True) for t in w]
        m = ShardedTwoLayerSage(build(lo, hi), x[lo:hi], lo, hi, *ws)
        m.world = group_world
        loss = m.loss(seeds, labels)
        loss.backward()
        m.sync_grads()
        return loss.detach(), [t.grad.clone() for t in ws], m.stats
    lo, hi = even_ranges(n, world)[rank]
    loss, grads, stats = run(lo, hi, world)
    ok = True
    if rank == 0:
        # single process, unsharded: the same class with world = 1 ...
        loss1, grads1, _ = run(0, n, 1)
        ok = torch.allclose(loss, loss1, rtol=1e-5) and all(torch.allclose(a, b, rtol=1e-4, atol=1e-6) for a, b in zip(grads, grads1))
        # ... and the oracle's restatement of the reference encoders stacked (src/graphsage.py:131-154, gcn=True)
        u1 = sorted(set(seeds.tolist()).union(*[adj[int(s)] for s in seeds]))
        h1 = oracle.sage_encoder(w[0], u1, adj, x, gcn=True).t()                    # [|U1|, h]
        pos = {v: i for i, v in enumerate(u1)}
        agg2 = torch.stack([h1[[pos[t] for t in sorted(adj[int(s)] | {int(s)})]].mean(0) for s in seeds])
        scores = torch.relu(agg2 @ w[1].t()) @ w[2].t()
        ref = torch.nn.functional.cross_entropy(scores, labels)
        ok = ok and torch.allclose(loss, ref, rtol=1e-5)
    allg = [torch.zeros_like(loss) for _ in range(world)]
    dist.all_gather(allg, loss)
    ok = ok and all(torch.equal(allg[0], t) for t in allg) and stats["u1"] > 60     # every rank ends with the same loss
    q.put((rank, bool(ok), float(loss)))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_feature_minibatch_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 500)
    procs = [ctx.Process(target=_sage_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] for r in res), res
