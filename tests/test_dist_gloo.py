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
