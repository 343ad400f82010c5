"""Multi-GPU layer pass: destination-node-range sharding, one collective per pass.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank g owns the CSR rows (destination
nodes) [lo_g, hi_g) of A and, for the backward, rows [lo'_g, hi'_g) of A^T; the dense operand is
replicated.  A pass is   local gather-reduce on the owned rows  ->  all-gather of the row shards.

    forward : Y[lo:hi]  = A[lo:hi, :] X           then all-gather(Y shards)  -> replicated Y
    backward: dX[lo':hi'] = A^T[lo':hi', :] dY    then all-gather(dX shards) -> replicated dX

which is the "one collective on the node accumulators per layer" of the north star with half the
traffic of a zero-padded all-reduce, no float atomics and a fixed summation order.  The exchange
can be chunked so that the all-gather of chunk k-1 (on a side stream) overlaps the gather-reduce of
chunk k.  The partitioner is pure integer work and is unit-tested on CPU (gloo).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def nnz_balanced_ranges(rowptr: np.ndarray, world: int, row_cost: float = 0.0) -> List[Tuple[int, int]]:
    """Contiguous row ranges with (as near as possible) equal work: exact integer prefix split of
    ``rowptr[i] + row_cost * i`` (row_cost is quantised to 1/16 so the arithmetic stays in int64).
    ``row_cost = 0`` balances nnz; ``row_cost = 1`` balances the merge-path items (rows + edges) of the
    gather kernel; larger values weight the output row (a full-width HBM write) against a gathered edge
    (which often hits L2) -- see fit_row_cost.  Deterministic and identical on every rank."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = len(rowptr) - 1
    q = int(round(float(row_cost) * 16))
    work = rowptr * 16 + q * np.arange(n + 1, dtype=np.int64) if q else rowptr
    total = int(work[-1])
    cuts = [0]
    for g in range(1, world):
        target = (total * g) // world
        r = int(np.searchsorted(work, target, side="left"))
        r = min(max(r, cuts[-1]), n)
        cuts.append(r)
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def fit_row_cost(stats: Sequence[Sequence[float]], lo: float = 0.5, hi: float = 8.0) -> Optional[float]:
    """Profile-guided weight for nnz_balanced_ranges: given one (rows, nnz, seconds) sample per rank of the
    same gather launch on differently shaped shards, least-squares fit  t = a * nnz + b * rows  and return
    b / a (cost of one output row in edges), clipped to [lo, hi]; None if the samples do not determine it."""
    m = np.asarray([[float(s[1]), float(s[0])] for s in stats], dtype=np.float64)
    t = np.asarray([float(s[2]) for s in stats], dtype=np.float64)
    if len(stats) < 2 or np.linalg.matrix_rank(m) < 2:
        return None
    (a, b), *_ = np.linalg.lstsq(m, t, rcond=None)
    if not (a > 0):
        return None
    return float(min(max(b / a, lo), hi))


def even_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    base, rem = divmod(n, world)
    out, lo = [], 0
    for g in range(world):
        hi = lo + base + (1 if g < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def all_gather_rows(local: torch.Tensor, ranges: Sequence[Tuple[int, int]], out: Optional[torch.Tensor] = None,
                    group=None) -> torch.Tensor:
    """Assemble the replicated [N, d] matrix from per-rank row shards (sizes may differ)."""
    world = len(ranges)
    n = ranges[-1][1]
    d = local.shape[1]
    if out is None:
        out = torch.empty(n, d, dtype=local.dtype, device=local.device)
    if world == 1:
        out.copy_(local)
        return out
    sizes = [hi - lo for lo, hi in ranges]
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    elif dist.get_backend(group) == "nccl":
        # NCCL handles ragged shards as one grouped broadcast, straight into the views of `out`
        dist.all_gather([out[lo:hi] for lo, hi in ranges], local.contiguous(), group=group)
    else:
        # gloo needs equal sizes: pad every shard to the largest one
        m = max(sizes)
        padded = torch.zeros(m, d, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        buf = torch.empty(world * m, d, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        for g, (lo, hi) in enumerate(ranges):
            out[lo:hi] = buf[g * m: g * m + (hi - lo)]
    return out


def halo_need_mask(consumer_col: torch.Tensor, ranges: Sequence[Tuple[int, int]], rank: int, group=None,
                   chunk: int = 1 << 26) -> torch.Tensor:
    """Which peers gather which of this rank's rows in their next pass (the halo of a layer exchange).

    ``consumer_col`` is the column array of the CSR shard THIS rank runs next on the replicated matrix (the
    backward's A^T rows, or the next layer's A rows); its distinct columns are the only rows of the
    replica this rank will ever read.  Every rank marks them, the marks are exchanged once, and the owner of
    rows [lo, hi) gets an int32 mask per owned row: bit s set <=> the s-th peer (ranks != ``rank`` in
    ascending order, the order of PeerReplica.peer_row_ptrs) needs that row.  Pure integer work, exact, and
    backend-agnostic (tested on CPU with gloo)."""
    world = len(ranges)
    n = ranges[-1][1]
    dev = consumer_col.device
    need = torch.zeros(n, dtype=torch.uint8, device=dev)
    one = torch.ones((), dtype=torch.uint8, device=dev)
    for part in consumer_col.split(chunk):
        need.index_put_((part.long(),), one)
    lo, hi = ranges[rank]
    mask = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
    if world == 1:
        return mask
    marks = [torch.empty_like(need) for _ in range(world)]
    dist.all_gather(marks, need, group=group)
    slot = 0
    for r in range(world):
        if r == rank:
            continue
        mask |= marks[r][lo:hi].to(torch.int32) << slot
        slot += 1
    return mask


class PeerReplica:
    """A replicated [N, d] fp32 matrix in CUDA symmetric memory (one copy per rank, peer-mapped over NVLink).

    Used for the *fused* exchange: the gather kernel's epilogue stores every finished row of the rank's
    shard into all replicas (``y_peers``: P2P stores) or once to the NVSwitch multicast address
    (``multicast``: multimem.st), so the all-gather is overlapped with the gather itself and no separate
    collective is launched.  ``barrier()`` (device-side, on the current stream) orders writers and readers."""

    def __init__(self, n_rows: int, d: int, ranges: Sequence[Tuple[int, int]], rank: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.ranges, self.d = rank, list(ranges), d
        self.buf = symm.empty((n_rows, d), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        try:
            self.multicast_ptr = int(self.hdl.multicast_ptr)
        except Exception:
            self.multicast_ptr = 0

    @property
    def local_rows(self) -> torch.Tensor:
        lo, hi = self.ranges[self.rank]
        return self.buf[lo:hi]

    @property
    def peer_row_ptrs(self) -> List[int]:
        lo, _ = self.ranges[self.rank]
        return [p + lo * self.d * 4 for r, p in enumerate(self.ptrs) if r != self.rank]

    @property
    def multicast_row_ptr(self) -> int:
        lo, _ = self.ranges[self.rank]
        return self.multicast_ptr + lo * self.d * 4 if self.multicast_ptr else 0

    def barrier(self, channel: int = 0) -> None:
        self.hdl.barrier(channel=channel)


class PeerBlock:
    """A per-rank [n_rows, d] fp32 block in CUDA symmetric memory that peers can write into with plain copies
    (``view(p)`` is a tensor aliasing rank p's block over NVLink).  Used to hand rows back to their owner, e.g.
    re-sharding the backward's dX from the compute-balanced source ranges to the owners' node ranges before it
    leaves the GPU."""

    def __init__(self, n_rows: int, d: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.shape = (n_rows, d)
        self.buf = symm.empty(self.shape, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)

    def view(self, rank: int) -> torch.Tensor:
        return self.hdl.get_buffer(rank, self.shape, torch.float32)

    def barrier(self, channel: int = 0) -> None:
        self.hdl.barrier(channel=channel)


def reshard_rows(local: torch.Tensor, src_range: Tuple[int, int], dst_ranges: Sequence[Tuple[int, int]],
                 block: "PeerBlock") -> None:
    """Copy this rank's rows [src_range) of a global row space into the owners' PeerBlocks (owner p holds
    rows dst_ranges[p]); one NVLink memcpy per overlapping owner.  Callers barrier before reading."""
    lo, hi = src_range
    for p, (a, b) in enumerate(dst_ranges):
        s, e = max(lo, a), min(hi, b)
        if s < e:
            block.view(p)[s - a:e - a].copy_(local[s - lo:e - lo], non_blocking=True)


class ShardedLayerPass:
    """Forward + backward of one aggregation layer on a node-range-sharded graph.

    ``fwd`` / ``bwd`` are this rank's CSRGraph shards (rows = owned destination / source range, columns =
    all nodes); ``compute`` is the local gather-reduce (ops.gather_reduce on the GPU; the CPU gloo test
    injects a stand-in)."""

    def __init__(self, fwd, bwd, fwd_ranges, bwd_ranges, rank: int, compute, chunks: int = 1, group=None):
        self.fwd, self.bwd = fwd, bwd
        self.fwd_ranges, self.bwd_ranges = list(fwd_ranges), list(bwd_ranges)
        self.rank, self.compute, self.chunks, self.group = rank, compute, max(1, chunks), group
        self.world = len(self.fwd_ranges)

    def _pass(self, g, ranges, x_full, out):
        local = self.compute(g, x_full)
        return all_gather_rows(local, ranges, out, self.group)

    def forward(self, x_full: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._pass(self.fwd, self.fwd_ranges, x_full, out)

    def backward(self, dy_full: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._pass(self.bwd, self.bwd_ranges, dy_full, out)
