"""Synthetic graphs of the BASELINE.json shapes (the real .mat / .npz datasets are not available).

* ``rmat_shard`` / ``rmat_transposed_shard``: the C5 / S64 generator of SURVEY.md 8(d) -- R-MAT
  (0.57, 0.19, 0.19, 0.05), generated ON DEVICE per destination shard from (seed, shard id), int32
  col / int64 rowptr, duplicates kept (multigraph; SpMM sums them like any CSR).
* ``planted_anomaly_graph``: small community graph with planted anomalies for AUROC parity runs
  (numpy, host).
* ``power_law_adj_lists``: dict-of-sets adjacency in the format program B un-pickles.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .graph import CSRGraph

RMAT_ABC = (0.57, 0.19, 0.19)


def _keys_to_csr(keys: torch.Tensor, n_keys: int, n_rows: int, n_cols: int, row_offset: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sort (row<<32|col) keys and split into rowptr / col.  Rows are global ids; ``row_offset`` is
    subtracted by searching for (row_offset + r) << 32."""
    if row_offset:
        keys[:n_keys] -= (row_offset << 32)
    rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=device)
    col = torch.empty(n_keys, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().ggad_coo_keys_to_csr(ptr(keys), n_keys, n_rows, ptr(rowptr), ptr(col), stream_ptr(device)))
        if n_keys >= (1 << 24):
            check(lib().ggad_trim_workspace())
    return rowptr, col


def rmat_shard(n_local: int, n_edges: int, n_shards: int = 1, shard: int = 0, seed: int = 0, device="cuda",
               mean: bool = True, abc=None) -> CSRGraph:
    """Destination shard ``shard``: rows = its n_local destination nodes, columns = all n_shards*n_local
    source nodes.  With ``mean`` the operator is the GraphSAGE mean aggregator (row_scale = 1/deg from
    exact integer degrees; empty rows give 0)."""
    device = torch.device(device)
    keys = torch.empty(n_edges, dtype=torch.int64, device=device)
    a, b, c = abc or RMAT_ABC
    with torch.cuda.device(device):
        check(lib().ggad_rmat_keys(ptr(keys), n_edges, n_local, n_shards, shard, seed, a, b, c, 0, 0, None,
                                   stream_ptr(device)))
    rowptr, col = _keys_to_csr(keys, n_edges, n_local, n_local * n_shards, shard * n_local, device)
    del keys
    g = CSRGraph(rowptr, col, None, n_local, n_local * n_shards)
    if mean:
        deg = g.degrees().to(torch.float32)
        g.row_scale = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg))
    return g


def rmat_transposed_shard(n_local: int, n_edges: int, n_shards: int, seed: int, lo: int, hi: int, device="cuda",
                          col_scale: Optional[torch.Tensor] = None, abc=None) -> CSRGraph:
    """Rows [lo, hi) of the TRANSPOSE of the whole n_shards-shard graph (rows = source nodes, columns =
    global destination ids), rebuilt by regenerating every shard's edges and keeping those whose source
    falls in the range.  ``col_scale`` (global, [n_shards*n_local]) carries the forward row scale."""
    device = torch.device(device)
    a, b, c = abc or RMAT_ABC
    scratch = torch.empty(n_edges, dtype=torch.int64, device=device)
    parts = []
    n_out = C.c_int64(0)
    for s in range(n_shards):
        with torch.cuda.device(device):
            check(lib().ggad_rmat_keys(ptr(scratch), n_edges, n_local, n_shards, s, seed, a, b, c, lo, hi,
                                       C.addressof(n_out), stream_ptr(device)))
        parts.append(scratch[: n_out.value].clone())
    keys = torch.cat(parts) if len(parts) > 1 else parts[0]
    del scratch, parts
    n_keys = int(keys.numel())
    rowptr, col = _keys_to_csr(keys, n_keys, hi - lo, n_local * n_shards, lo, device)
    del keys
    return CSRGraph(rowptr, col, None, hi - lo, n_local * n_shards, col_scale=col_scale)


# ------------------------------------------------------------------------------------------
def planted_anomaly_graph(n: int, avg_deg: float, d: int, anomaly_rate: float, seed: int = 0, n_comm: int = 8):
    """Symmetric binary community graph with planted anomalies.  Normal nodes: features = community
    centroid + noise, edges mostly inside the community.  Anomalies: features from a mixture of foreign
    centroids and half of their edges rewired at random (lower local affinity).  Returns
    (scipy CSR adjacency, features [n,d] fp32, labels [n] int64)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    comm = rng.integers(0, n_comm, n)
    cent = rng.standard_normal((n_comm, d)).astype(np.float32)
    labels = (rng.random(n) < anomaly_rate).astype(np.int64)
    x = cent[comm] + 0.35 * rng.standard_normal((n, d)).astype(np.float32)
    ab = np.flatnonzero(labels)
    x[ab] = 0.5 * (cent[rng.integers(0, n_comm, len(ab))] + cent[rng.integers(0, n_comm, len(ab))]) \
        + 0.9 * rng.standard_normal((len(ab), d)).astype(np.float32)
    m = int(n * avg_deg / 2)
    src = rng.integers(0, n, m)
    order = np.argsort(comm, kind="stable")
    start = np.searchsorted(comm[order], np.arange(n_comm))
    size = np.bincount(comm, minlength=n_comm)
    same = rng.random(m) < 0.9
    pick = (rng.random(m) * size[comm[src]]).astype(np.int64)
    dst_same = order[np.minimum(start[comm[src]] + pick, n - 1)]
    dst = np.where(same & (labels[src] == 0), dst_same, rng.integers(0, n, m))
    keep = src != dst
    a = sp.coo_matrix((np.ones(keep.sum()), (src[keep], dst[keep])), shape=(n, n)).tocsr()
    a = ((a + a.T) > 0).astype(np.float64).tocsr()
    return a, x.astype(np.float32), labels


def power_law_adj_lists(n: int, avg_deg: float, seed: int = 0, alpha: float = 2.1, max_deg: int = 100000):
    """dict[int -> set[int]] (symmetric, no self loops) with a truncated power-law degree profile."""
    from collections import defaultdict
    rng = np.random.default_rng(seed)
    w = (1.0 - rng.random(n)) ** (-1.0 / (alpha - 1.0))
    w = np.minimum(w, max_deg)
    p = w / w.sum()
    m = int(n * avg_deg / 2)
    src = rng.choice(n, m, p=p)
    dst = rng.integers(0, n, m)
    adj = defaultdict(set)
    for s, t in zip(src.tolist(), dst.tolist()):
        if s != t:
            adj[s].add(t)
            adj[t].add(s)
    for v in range(n):
        if len(adj[v]) == 0:
            t = (v + 1) % n
            adj[v].add(t)
            adj[t].add(v)
    return adj


def rmat_adjacency(n: int, n_edges: int, seed: int = 0, device="cuda"):
    """Symmetric, duplicate-free, self-loop-free R-MAT adjacency as a DeviceAdjacency (sorted neighbor lists) --
    the DGraph-shaped input of program B without going through a Python dict of sets.  ``n_edges`` directed
    candidates are generated; the stored entry count (both directions, after de-duplication) is returned by
    ``adj.col.numel()``."""
    from .graph import DeviceAdjacency
    device = torch.device(device)
    keys = torch.empty(n_edges, dtype=torch.int64, device=device)
    a, b, c = RMAT_ABC
    with torch.cuda.device(device):
        check(lib().ggad_rmat_keys(ptr(keys), n_edges, n, 1, 0, seed, a, b, c, 0, 0, None, stream_ptr(device)))
    dst, src = keys >> 32, keys & 0xffffffff
    keep = dst != src
    dst, src = dst[keep], src[keep]
    both = torch.cat([(dst << 32) | src, (src << 32) | dst])
    del keys, dst, src, keep
    both = torch.unique(both)                        # sorted + de-duplicated
    m = int(both.numel())
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=device)
    col = torch.empty(m, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().ggad_coo_keys_to_csr(ptr(both), m, n, ptr(rowptr), ptr(col), stream_ptr(device)))
        if m >= (1 << 24):
            check(lib().ggad_trim_workspace())
    return DeviceAdjacency(rowptr, col, n)
