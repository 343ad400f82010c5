"""Mini-batch aggregation on a node-range-SHARDED feature table: partial accumulators + one all-reduce per layer
(third bullet; BASELINE.json config 5: "mini-batch SAGE two-layer, 8xB200 NCCL all-reduce").

The reference has no multi-GPU code; this is the scheme for graphs whose feature table should not be replicated
(C5: 50 M nodes x 64 floats = 12.8 GB per replica, 1.6 GB per rank when sharded).  Rank g keeps

    x_local = X[lo_g:hi_g]                          its slice of the frozen feature table
    adj_g   = A[:, lo_g:hi_g]                       every node's neighbor list restricted to its column range
                                                    (CSR over ALL rows, LOCAL column ids)

and all ranks process the SAME super-batch of seeds S:

    hop 1   U1 = S  u  union_g N_g(S)               id lists all-gathered (integers, exact)
    layer 1 part1_g[u] = sum_{v in N_g(u)} x_v (+ x_u on u's owner)     ONE gather-reduce over x_local, no remap
            deg(u)   = sum_g |N_g(u)|                int all-reduce (exact), normalisation AFTER the float reduce
            agg1     = all_reduce(part1) / (deg + 1) ;   h1 = ReLU(agg1 W1^T)        [|U1|, d] floats on the wire
    layer 2 part2_g[s] = sum_{u in N_g(s)} h1[u] (+ h1[s] on s's owner) ;  agg2 = all_reduce(part2) / (deg + 1)
            h2 = ReLU(agg2 W2^T) ; scores = h2 Wc^T                                  [|S|, h] floats on the wire

which is MeanAggregator / Encoder with gcn=True, num_sample=None stacked twice (src/graphsage.py:66-99,131-154 and
the graphsage-simple idiom of :108-121).  Every rank ends with the same loss.  Backward: the all-reduce is the
identity for the gradient (each rank back-propagates through ITS part of the hop-1 block), so the gradient of W1 is
a partial sum that ``sync_grads`` all-reduces (h x d floats); W2 / Wc gradients are already complete and identical.
fp32 summation order differs from the single-GPU pass -> compare at rtol 1e-5.

The compute primitives are injected (``backend``) so that the index / collective logic is unit-tested on CPU with
gloo (tests/test_dist_gloo.py); the product backend below runs the CUDA kernels only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


class DeviceBackend:
    """CUDA primitives: block extraction (ggad_block_rowptr / _fill), gather-reduce (ops.spmm), projection
    (ops.linear)."""

    def __init__(self, adj_shard):
        self.adj = adj_shard                     # graph.DeviceAdjacency: rows = all nodes, cols = LOCAL ids

    def block(self, nodes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        import ctypes as C
        from ._lib import check, lib, ptr, stream_ptr
        a, dev = self.adj, self.adj.device
        nodes = nodes.to(device=dev, dtype=torch.int32).contiguous()
        nb = int(nodes.numel())
        rowptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib().ggad_block_rowptr(ptr(a.rowptr), ptr(a.col), a.n, ptr(nodes), nb, 0, ptr(rowptr), C.addressof(nnz), st))
            cols = torch.empty(max(int(nnz.value), 1), dtype=torch.int32, device=dev)
            check(lib().ggad_block_fill(ptr(a.rowptr), ptr(a.col), a.n, ptr(nodes), nb, 0, ptr(rowptr), ptr(cols), st))
        return rowptr, cols[: int(nnz.value)]

    def spmm(self, rowptr, col, n_rows, n_cols, table):
        from . import ops
        from .graph import CSRGraph
        return ops.spmm(CSRGraph(rowptr, col.to(torch.int32), None, n_rows, n_cols), table)

    def linear(self, x, w, relu=False):
        from . import ops
        return ops.linear(x, w, relu=relu)


class _AllReduceSum(torch.autograd.Function):
    """sum over ranks; backward = identity (every rank holds the same downstream graph, see module docstring)."""

    @staticmethod
    def forward(ctx, t, group):
        out = t.contiguous().clone()
        dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_gather_ids(ids: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenation of every rank's (variable-length) int64 id list; works with gloo and nccl."""
    world = dist.get_world_size(group)
    n = torch.tensor([ids.numel()], dtype=torch.int64, device=ids.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=torch.int64, device=ids.device)
    pad[: ids.numel()] = ids
    buf = torch.empty(world * m, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)])


class ShardedTwoLayerSage:
    """Two stacked mean-SAGE layers (gcn=True) + a linear classifier on a column-range-sharded graph."""

    def __init__(self, backend, x_local: torch.Tensor, lo: int, hi: int, w1: torch.Tensor, w2: torch.Tensor,
                 w_cls: torch.Tensor, group=None):
        self.be, self.x, self.lo, self.hi, self.group = backend, x_local, int(lo), int(hi), group
        self.w1, self.w2, self.w_cls = w1, w2, w_cls
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stats = {}

    def _reduce(self, t: torch.Tensor, differentiable: bool = False) -> torch.Tensor:
        if self.world == 1:
            return t
        if differentiable:
            return _AllReduceSum.apply(t, self.group)
        t = t.contiguous()
        dist.all_reduce(t, group=self.group)
        return t

    def forward(self, seeds: torch.Tensor) -> torch.Tensor:
        """scores [|S|, n_classes] for the int64 seed ids (identical on every rank)."""
        be, lo, hi = self.be, self.lo, self.hi
        dev = self.x.device
        seeds = seeds.to(dev, torch.int64)
        # ---- hop 1: local neighbor lists of the seeds, global frontier by all-gather of the id lists ----
        rp1, c1 = be.block(seeds)
        g1 = c1.to(torch.int64) + lo
        cand = torch.unique(g1)
        if self.world > 1:
            cand = all_gather_ids(cand, self.group)
        u1 = torch.unique(torch.cat([cand, seeds]))                                  # sorted, identical on every rank
        deg_s = self._reduce((rp1[1:] - rp1[:-1]).clone())                           # exact integer degrees
        # ---- layer 1: partial sums over the locally owned neighbor rows, ONE float all-reduce ----
        rp2, c2 = be.block(u1)
        deg_u = self._reduce((rp2[1:] - rp2[:-1]).clone())
        part1 = be.spmm(rp2, c2, int(u1.numel()), int(self.x.shape[0]), self.x)
        own_u = (u1 >= lo) & (u1 < hi)
        part1 = part1.index_add(0, torch.nonzero(own_u).reshape(-1), self.x[(u1[own_u] - lo)])   # self term, owner only
        agg1 = self._reduce(part1) / (deg_u + 1).to(torch.float32).unsqueeze(1)
        h1 = be.linear(agg1, self.w1, relu=True)                                     # [|U1|, h], replicated
        # ---- layer 2: the hop-1 block again, columns remapped into U1; partial sums over h1, ONE all-reduce ----
        idx = torch.searchsorted(u1, g1)
        part2 = be.spmm(rp1, idx, int(seeds.numel()), int(u1.numel()), h1)
        own_s = (seeds >= lo) & (seeds < hi)
        pos_s = torch.searchsorted(u1, seeds)
        part2 = part2.index_add(0, torch.nonzero(own_s).reshape(-1), h1[pos_s[own_s]])
        agg2 = self._reduce(part2, differentiable=True) / (deg_s + 1).to(torch.float32).unsqueeze(1)
        h2 = be.linear(agg2, self.w2, relu=True)
        self.stats = dict(u1=int(u1.numel()), hop1_edges=int(c1.numel()), hop2_edges=int(c2.numel()),
                          wire_floats=int(u1.numel()) * self.x.shape[1] + int(seeds.numel()) * h1.shape[1])
        return be.linear(h2, self.w_cls)

    def loss(self, seeds: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """GraphSage.loss (src/graphsage.py:39-43): cross entropy of the scores."""
        return torch.nn.functional.cross_entropy(self.forward(seeds), labels.to(self.x.device).reshape(-1))

    def sync_grads(self) -> None:
        """Complete the partial gradient of W1 (see module docstring); W2 / W_cls gradients are already complete."""
        if self.world > 1 and self.w1.grad is not None:
            dist.all_reduce(self.w1.grad, group=self.group)


def column_shard(adj, lo: int, hi: int):
    """A[:, lo:hi] of a device adjacency (graph.DeviceAdjacency, global sorted neighbor ids) as a DeviceAdjacency over
    ALL rows with LOCAL column ids -- the per-rank piece of the sharded scheme.  Integer work, exact."""
    from .graph import DeviceAdjacency
    keep = (adj.col >= lo) & (adj.col < hi)
    csum = torch.zeros(adj.col.numel() + 1, dtype=torch.int64, device=adj.device)
    torch.cumsum(keep, 0, out=csum[1:])
    return DeviceAdjacency(csum[adj.rowptr].contiguous(), (adj.col[keep] - lo).to(torch.int32).contiguous(), adj.n)
