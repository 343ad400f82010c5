"""Device-resident CSR container for the aggregation operators of the GGAD hot path.

HBM layout: rowptr int64 [n_rows+1], col int32 [nnz], val fp32 [nnz] or absent (all ones, with the
normalisation carried by row_scale / col_scale vectors), plus the merge-path plan (tile_row int32,
tile_edge int64, one entry per 2048 items) and a per-width partial-row workspace.  The transpose
(needed by every backward) is built lazily on the device and cached.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

# below this many (rows + edges) -- one merge-path tile -- the group-per-row kernel is used and no plan is built.
# Anything larger goes through the tiled kernel: a mini-batch block of a few thousand edges can hold one hub row with
# thousands of neighbors, which the group-per-row kernel walks serially (measured: 0.65 ms for an 8 k-edge hop-1 block)
PLAN_MIN_ITEMS = 2048


def _pad4(n: int) -> int:
    return (n + 3) & ~3


TRIM_AFTER_NNZ = 1 << 24   # sorts of at least this many keys release the library's temp pool afterwards


class CSRGraph:
    """A sparse operator A [n_rows, n_cols] in CSR on one GPU.

    Effective entry: A[r, c] = row_scale[r] * val[e] * col_scale[c]  (missing parts = 1).
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], n_rows: int, n_cols: int,
                 row_scale: Optional[torch.Tensor] = None, col_scale: Optional[torch.Tensor] = None,
                 symmetric_pattern: bool = False, use_plan: Optional[bool] = None):
        for name, t in (("rowptr", rowptr), ("col", col)):
            _lib.require_cuda(t, name)
        assert rowptr.dtype == torch.int64 and col.dtype == torch.int32
        assert val is None or (val.dtype == torch.float32 and val.is_cuda)
        self.rowptr, self.col, self.val = rowptr.contiguous(), col.contiguous(), None if val is None else val.contiguous()
        self.n_rows, self.n_cols, self.nnz = int(n_rows), int(n_cols), int(col.numel())
        assert self.rowptr.numel() == self.n_rows + 1
        self.row_scale, self.col_scale = row_scale, col_scale
        self.device = rowptr.device
        # pattern AND values symmetric (A == A^T up to the row/col scale swap): transpose reuses the arrays
        self.symmetric_pattern = symmetric_pattern
        self._use_plan = (self.n_rows + self.nnz >= PLAN_MIN_ITEMS) if use_plan is None else use_plan
        self._plan = None
        self._ws: Dict[int, torch.Tensor] = {}
        self._T: Optional["CSRGraph"] = None
        self._rows_cache: Dict[bytes, "CSRGraph"] = {}

    # ---------------------------------------------------------------- builders
    @classmethod
    def from_arrays(cls, rowptr, col, val, n_rows, n_cols, device="cuda", **kw) -> "CSRGraph":
        rp = torch.as_tensor(np.asarray(rowptr, dtype=np.int64)).to(device)
        c = torch.as_tensor(np.asarray(col, dtype=np.int32)).to(device)
        v = None if val is None else torch.as_tensor(np.asarray(val, dtype=np.float32)).to(device)
        return cls(rp, c, v, n_rows, n_cols, **kw)

    @classmethod
    def from_scipy(cls, m, device="cuda", **kw) -> "CSRGraph":
        import scipy.sparse as sp
        m = sp.csr_matrix(m)
        m.sum_duplicates()
        m.sort_indices()
        return cls.from_arrays(m.indptr, m.indices, m.data, m.shape[0], m.shape[1], device=device, **kw)

    @classmethod
    def from_dense(cls, adj: torch.Tensor, device="cuda", **kw) -> "CSRGraph":
        """From the dense fp32 ``adj`` / ``raw_adj`` tensors run.py builds ([1,N,N] or [N,N])."""
        a = adj.detach()
        if a.dim() == 3:
            a = a[0]
        sp_t = a.to_sparse_csr()
        return cls(sp_t.crow_indices().to(torch.int64).to(device), sp_t.col_indices().to(torch.int32).to(device),
                   sp_t.values().to(torch.float32).to(device), a.shape[0], a.shape[1], **kw)

    @classmethod
    def from_any(cls, adj, device="cuda") -> "CSRGraph":
        if isinstance(adj, CSRGraph):
            return adj
        if isinstance(adj, torch.Tensor):
            if adj.layout == torch.strided:
                return cls.from_dense(adj, device)
            a = adj.coalesce() if adj.layout == torch.sparse_coo else adj
            if a.layout == torch.sparse_coo:
                a = a.to_sparse_csr()
            return cls(a.crow_indices().to(torch.int64).to(device), a.col_indices().to(torch.int32).to(device),
                       a.values().to(torch.float32).to(device), a.shape[-2], a.shape[-1])
        return cls.from_scipy(adj, device)

    # ---------------------------------------------------------------- plan / workspace
    @property
    def plan(self):
        """(tile_row, tile_edge, n_tiles) or None for small graphs."""
        if not self._use_plan:
            return None
        if self._plan is None:
            n_tiles = int(lib().ggad_plan_num_tiles(self.n_rows, self.nnz))
            tile_row = torch.empty(n_tiles + 1, dtype=torch.int32, device=self.device)
            tile_edge = torch.empty(n_tiles + 1, dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                check(lib().ggad_plan_build(ptr(self.rowptr), self.n_rows, self.nnz, ptr(tile_row), ptr(tile_edge),
                                            stream_ptr(self.device)))
            self._plan = (tile_row, tile_edge, n_tiles)
        return self._plan

    def chase_flags(self):
        """(tile_done int32[n_tiles], epoch) for the chase exchange: a fresh epoch per launch, so the flags are never
        reset (the buffer starts at 0, epochs start at 1)."""
        if self.__dict__.get("_done") is None:
            self._done = torch.zeros(self.plan[2], dtype=torch.int32, device=self.device)
            self._epoch = 0
        self._epoch = self._epoch % 0x7ffffff0 + 1
        return self._done, self._epoch

    def workspace(self, d: int) -> Optional[torch.Tensor]:
        if self.plan is None:
            return None
        ws = self._ws.get(d)
        if ws is None:
            ws = torch.empty(2 * self.plan[2] * d, dtype=torch.float32, device=self.device)
            self._ws = {d: ws}          # keep only the latest width
        return ws

    # ---------------------------------------------------------------- derived operators
    @property
    def T(self) -> "CSRGraph":
        """Transposed operator (device build, cached).  Row/col scales swap roles."""
        if self._T is None:
            if self.symmetric_pattern:
                t = CSRGraph(self.rowptr, self.col, self.val, self.n_cols, self.n_rows, row_scale=self.col_scale,
                             col_scale=self.row_scale, symmetric_pattern=True, use_plan=self._use_plan)
                t._plan = self._plan
            else:
                rpt = torch.empty(self.n_cols + 1, dtype=torch.int64, device=self.device)
                ct = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
                vt = None if self.val is None else torch.empty(self.nnz, dtype=torch.float32, device=self.device)
                with torch.cuda.device(self.device):
                    check(lib().ggad_csr_transpose(ptr(self.rowptr), ptr(self.col), ptr(self.val), self.n_rows, self.n_cols,
                                                   self.nnz, ptr(rpt), ptr(ct), ptr(vt), None, stream_ptr(self.device)))
                    if self.nnz >= TRIM_AFTER_NNZ:      # one-off multi-GB sort buffers: hand them back to the driver
                        check(lib().ggad_trim_workspace())
                t = CSRGraph(rpt, ct, vt, self.n_cols, self.n_rows, row_scale=self.col_scale, col_scale=self.row_scale)
                t = t.fold_col_scale()
            t._T = self
            self._T = t
        return self._T

    def rows(self, idx: Sequence[int]) -> "CSRGraph":
        """Sub-operator A[idx, :] as its own CSR (the ``adj[0, S, :]`` of model.py:151), cached per index set."""
        idx_np = np.asarray(idx, dtype=np.int32)
        key = idx_np.tobytes()
        g = self._rows_cache.get(key)
        if g is None:
            rows = torch.from_numpy(idx_np).to(self.device)
            deg = (self.rowptr[1:] - self.rowptr[:-1])[rows.long()]
            sub_ptr = torch.zeros(len(idx_np) + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(deg, 0, out=sub_ptr[1:])
            nnz = int(sub_ptr[-1].item())
            sub_col = torch.empty(nnz, dtype=torch.int32, device=self.device)
            sub_val = None if self.val is None else torch.empty(nnz, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                check(lib().ggad_csr_extract_rows(ptr(self.rowptr), ptr(self.col), ptr(self.val), ptr(rows), len(idx_np),
                                                  ptr(sub_ptr), ptr(sub_col), ptr(sub_val), stream_ptr(self.device)))
            rs = None if self.row_scale is None else self.row_scale[rows.long()].contiguous()
            g = CSRGraph(sub_ptr, sub_col, sub_val, len(idx_np), self.n_cols, row_scale=rs, col_scale=self.col_scale)
            if len(self._rows_cache) > 8:
                self._rows_cache.clear()
            self._rows_cache[key] = g
        return g

    def fold_col_scale(self) -> "CSRGraph":
        """Same operator with the column scale folded into per-edge values (val[e] *= col_scale[col[e]]).
        A streamed 4 B/edge value is far cheaper than a random 4 B gather (one 32 B sector) per edge."""
        if self.col_scale is None:
            return self
        v = self.col_scale[self.col.long()]
        if self.val is not None:
            v = v * self.val
        g = CSRGraph(self.rowptr, self.col, v.contiguous(), self.n_rows, self.n_cols, row_scale=self.row_scale,
                     col_scale=None, symmetric_pattern=False, use_plan=self._use_plan)
        g._plan = self._plan
        return g

    def reorder_edges_hot_first(self, col_weight: Optional[torch.Tensor] = None, serpentine: bool = False) -> "CSRGraph":
        """Same operator with the edges of every row re-ordered by the popularity of their column (most gathered
        column first; ``col_weight`` defaults to the exact integer in-degree histogram of ``col``).  Pure index work,
        done once per graph on the device.  Why: the kernel gathers U rows per lane group at a time and a batch
        completes at the latency of its slowest row; popular columns are the ones resident in L2, so this order makes
        the batches homogeneous -- all-hit batches finish at L2 latency instead of waiting for one DRAM miss.  The
        result differs from the column-sorted order only in the fp32 summation order (still deterministic).
        ``serpentine``: odd rows are ordered coldest-first, so that consecutive rows meet hot-to-hot and cold-to-cold
        (rows are shorter than a few batches; a lane group walks them back to back)."""
        if self.nnz == 0:
            return self
        dev = self.device
        w = torch.bincount(self.col, minlength=self.n_cols) if col_weight is None else col_weight.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=dev), self.degrees())
        top = int(w.max().item()) + 1
        wc = w[self.col.long()]
        key = rows * top + (torch.where((rows & 1) == 1, wc, top - 1 - wc) if serpentine else (top - 1 - wc))
        del wc
        del rows
        perm = torch.argsort(key)
        del key
        g = CSRGraph(self.rowptr, self.col[perm].contiguous(), None if self.val is None else self.val[perm].contiguous(),
                     self.n_rows, self.n_cols, row_scale=self.row_scale, col_scale=self.col_scale,
                     symmetric_pattern=False, use_plan=self._use_plan)
        g._plan = self._plan                     # the plan depends on rowptr only
        return g

    def degrees(self) -> torch.Tensor:
        """Exact integer row degrees (int64)."""
        return self.rowptr[1:] - self.rowptr[:-1]

    def to_scipy(self):
        import scipy.sparse as sp
        v = np.ones(self.nnz, dtype=np.float32) if self.val is None else self.val.cpu().numpy()
        m = sp.csr_matrix((v, self.col.cpu().numpy(), self.rowptr.cpu().numpy()), shape=(self.n_rows, self.n_cols))
        if self.row_scale is not None:
            m = sp.diags(self.row_scale.cpu().numpy()).dot(m)
        if self.col_scale is not None:
            m = m.dot(sp.diags(self.col_scale.cpu().numpy()))
        return sp.csr_matrix(m)

    def algorithmic_bytes(self, d: int) -> int:
        """B_alg of SURVEY.md 8(d): every input read once, output written once."""
        w = 0 if self.val is None else 1
        return self.nnz * (4 + 4 * w) + (self.n_rows + 1) * 8 + self.n_cols * d * 4 + self.n_rows * d * 4


# ------------------------------------------------------------------------------------------
# host-side index work (bit-exact restatement of the reference preprocessing)
# ------------------------------------------------------------------------------------------
def normalize_adj_scipy(adj):
    """``D^-1/2 A^T D^-1/2`` in fp64 with D from row sums -- same scipy calls as utils.py:47-54."""
    import scipy.sparse as sp
    a = sp.coo_matrix(adj)
    deg = np.asarray(a.sum(1), dtype=np.float64).reshape(-1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)
    dis[np.isinf(dis)] = 0.0
    dm = sp.diags(dis)
    return a.dot(dm).transpose().dot(dm).tocsr()


def full_batch_graphs(adj, device="cuda"):
    """(A_hat, R, R^T) CSRGraphs for program A from the raw scipy adjacency, as run.py:96-109:
    A_hat = normalize_adj(A) + I, R = A + I (fp64, rounded once to fp32)."""
    import scipy.sparse as sp
    a = sp.csr_matrix(adj).astype(np.float64)
    n = a.shape[0]
    eye = sp.eye(n, dtype=np.float64, format="csr")
    a_hat = (normalize_adj_scipy(a) + eye).tocsr()
    r = (a + eye).tocsr()
    sym = (abs(a - a.T)).nnz == 0
    g_hat = CSRGraph.from_scipy(a_hat.astype(np.float32), device, symmetric_pattern=sym)
    g_r = CSRGraph.from_scipy(r.astype(np.float32), device, symmetric_pattern=sym)
    return g_hat, g_r


class AdjListCSR:
    """Host CSR view of the ``dict[int -> set[int]]`` adjacency lists of program B
    (src/utils.py:27-28,96-112), built once; neighbor ids sorted."""

    _cache: Dict[int, tuple] = {}

    def __init__(self, adj_lists, n: Optional[int] = None):
        keys = np.fromiter((int(k) for k in adj_lists.keys()), dtype=np.int64)
        n_nodes = int(max(keys.max() + 1 if len(keys) else 0, n or 0))
        for v in adj_lists.values():
            if len(v):
                n_nodes = max(n_nodes, int(max(v)) + 1)
        rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
        for k, v in adj_lists.items():
            rowptr[int(k) + 1] = len(v)
        np.cumsum(rowptr, out=rowptr)
        col = np.empty(rowptr[-1], dtype=np.int64)
        for k, v in adj_lists.items():
            k = int(k)
            if len(v):
                col[rowptr[k]:rowptr[k + 1]] = np.sort(np.fromiter((int(t) for t in v), dtype=np.int64, count=len(v)))
        self.rowptr, self.col, self.n = rowptr, col, n_nodes
        self.n_keys = len(adj_lists)

    def device(self, device) -> "DeviceAdjacency":
        """Device copy of the adjacency CSR (uploaded once per device)."""
        key = str(device)
        cache = self.__dict__.setdefault("_dev", {})
        if key not in cache:
            cache[key] = DeviceAdjacency(torch.from_numpy(self.rowptr).to(device),
                                         torch.from_numpy(self.col.astype(np.int32)).to(device), self.n)
        return cache[key]

    @classmethod
    def get(cls, adj_lists) -> "AdjListCSR":
        """Cached conversion of a dict-of-sets adjacency (the cache keeps a reference to the dict)."""
        key = id(adj_lists)
        hit = cls._cache.get(key)
        if hit is None or hit[0] is not adj_lists or hit[1].n_keys != len(adj_lists):
            hit = (adj_lists, cls(adj_lists))
            if len(cls._cache) > 4:
                cls._cache.clear()
            cls._cache[key] = hit
        return hit[1]

    def neighbors(self, nodes: np.ndarray, add_self: bool):
        """Block (rows=len(nodes)) -> (rowptr, col_global) with per-row sorted unique neighbor ids
        (optionally united with the node itself, set semantics)."""
        nodes = np.asarray(nodes, dtype=np.int64)
        safe = np.minimum(nodes, max(self.n - 1, 0))
        inside = nodes < self.n
        s = np.where(inside, self.rowptr[safe], 0)
        e = np.where(inside, self.rowptr[safe + 1], 0)
        lens = e - s
        total = int(lens.sum())
        out_ptr = np.zeros(len(nodes) + 1, dtype=np.int64)
        np.cumsum(lens, out=out_ptr[1:])
        offs = np.repeat(s - out_ptr[:-1], lens) + np.arange(total, dtype=np.int64)   # flat gather of the slices
        cols = self.col[offs]
        if add_self and len(nodes):
            rows = np.concatenate([np.repeat(np.arange(len(nodes), dtype=np.int64), lens),
                                   np.arange(len(nodes), dtype=np.int64)])
            cols = np.concatenate([cols, nodes])
            m = int(max(self.n, nodes.max() + 1))
            key = np.unique(rows * m + cols)                 # drops a duplicate self entry, sorts (row, col)
            rows, cols = key // m, key % m
            out_ptr = np.zeros(len(nodes) + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=len(nodes)), out=out_ptr[1:])
        return out_ptr, cols


class DeviceAdjacency:
    """Adjacency lists as a device CSR (rowptr int64, col int32 with sorted neighbor ids) -- the input of the
    device-side frontier construction (ggad_block_* entry points)."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, n: int):
        _lib.require_cuda(rowptr, "rowptr")
        assert rowptr.dtype == torch.int64 and col.dtype == torch.int32
        self.rowptr, self.col, self.n, self.device = rowptr.contiguous(), col.contiguous(), int(n), rowptr.device
        self._reserved_m = 0
        self.reserve_edges = 0      # optional hint: expected upper bound of a hop's block edges (reserved at the first batch)

    def _reserve(self, m: int, st) -> None:
        """Frontier sizes vary a lot from batch to batch (hubs).  At every new maximum, grow both memory pools ONCE
        to twice that size -- the library's pool for the sort buffers and torch's caching allocator for the four
        block arrays -- so that steady-state batches never reach cudaMalloc / cuMemCreate (which cost milliseconds,
        tens of milliseconds when the memory is peer-mapped to 7 other GPUs, and stall every rank of a
        data-parallel step)."""
        if m <= self._reserved_m:
            return
        self._reserved_m = 2 * m
        check(lib().ggad_reserve_workspace(2 * (8 * self._reserved_m + (32 << 20)), st))
        prime = [torch.empty(self._reserved_m, dtype=torch.int32, device=self.device) for _ in range(4)]
        del prime

    def block(self, nodes: torch.Tensor, add_self: bool):
        """One aggregation hop for ``nodes`` (int32 device tensor): the union frontier and the block CSR with
        exact integer degrees, all on the device.  Two host round trips (block nnz and |frontier|).
        Returns dict(frontier, rowptr, col, cdeg, n_rows, n_cols) of device tensors / ints."""
        import ctypes as C
        dev = self.device
        nodes = nodes.to(device=dev, dtype=torch.int32).contiguous()
        nb = int(nodes.numel())
        block_rowptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        h = lib()
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(h.ggad_block_rowptr(ptr(self.rowptr), ptr(self.col), self.n, ptr(nodes), nb, int(add_self),
                                      ptr(block_rowptr), C.addressof(nnz), st))
            m = int(nnz.value)
            self._reserve(max(m, int(self.reserve_edges)), st)
            cols = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
            check(h.ggad_block_fill(ptr(self.rowptr), ptr(self.col), self.n, ptr(nodes), nb, int(add_self),
                                    ptr(block_rowptr), ptr(cols), st))
            uniq = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
            nu = C.c_int64(0)
            check(h.ggad_unique_sorted(ptr(cols), m, 1 << 31, ptr(uniq), C.addressof(nu), st))
            k = int(nu.value)
            local = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
            cdeg = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
            check(h.ggad_block_remap(ptr(cols), m, ptr(uniq), k, ptr(local), ptr(cdeg), st))
        return dict(frontier=uniq[:k], rowptr=block_rowptr, col=local[:m], cdeg=cdeg[:k], n_rows=nb, n_cols=k)


def _block_direct(self, nodes: torch.Tensor, add_self: bool):
    """Hop block for a gather straight from the feature table: rows = ``nodes``, columns = GLOBAL neighbor ids, per-edge
    value 1/sqrt(batch-local column degree) from an exact integer histogram (ggad_block_col_weights).  No frontier list,
    no local ids: no sort / unique / remap and one size read-back instead of two."""
    import ctypes as C
    dev = self.device
    nodes = nodes.to(device=dev, dtype=torch.int32).contiguous()
    nb = int(nodes.numel())
    block_rowptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    nnz = C.c_int64(0)
    h = lib()
    with torch.cuda.device(dev):
        st = stream_ptr(dev)
        check(h.ggad_block_rowptr(ptr(self.rowptr), ptr(self.col), self.n, ptr(nodes), nb, int(add_self),
                                  ptr(block_rowptr), C.addressof(nnz), st))
        m = int(nnz.value)
        self._reserve(max(m, int(self.reserve_edges)), st)          # allocator priming at new maxima (see _reserve)
        cols = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
        check(h.ggad_block_fill(ptr(self.rowptr), ptr(self.col), self.n, ptr(nodes), nb, int(add_self),
                                ptr(block_rowptr), ptr(cols), st))
        # one histogram scratch per stream (the prefetch thread builds blocks on its own stream)
        key = torch.cuda.current_stream(dev).cuda_stream
        counts = self.__dict__.setdefault("_counts", {}).get(key)
        if counts is None:
            counts = torch.empty(self.n, dtype=torch.int32, device=dev)
            self._counts[key] = counts
        val = torch.empty(max(m, 1), dtype=torch.float32, device=dev)
        check(h.ggad_block_col_weights(ptr(cols), m, ptr(counts), self.n, ptr(val), st))
    return dict(rowptr=block_rowptr, col=cols[:m], val=val[:m], n_rows=nb, n_cols=self.n)


DeviceAdjacency.block_direct = _block_direct


def batch_block(adj: AdjListCSR, nodes: Sequence[int], add_self: bool):
    """Frontier + local block CSR for one aggregation hop (the set unions of
    src/graphsage.py:305-311,335-341 as array ops).  Returns dict with:
    frontier (sorted unique global ids), rowptr, col_local (int32), rdeg, cdeg (exact ints)."""
    rowptr, cols = adj.neighbors(np.asarray(nodes, dtype=np.int64), add_self)
    frontier, col_local = np.unique(cols, return_inverse=True)
    rdeg = np.diff(rowptr)
    cdeg = np.bincount(col_local, minlength=len(frontier)).astype(np.int64)
    return dict(frontier=frontier, rowptr=rowptr, col=col_local.astype(np.int32), rdeg=rdeg, cdeg=cdeg)


def full_batch_graphs_device(adj, device="cuda"):
    """(A_hat, R) like ``full_batch_graphs`` but with the preprocessing ON THE DEVICE (utils.py:47-54, run.py:98-101):
    fp64 row sums, D^-1/2, the transpose, the two fp64 products in scipy's order, "+ I" and the single rounding to
    fp32 are CUDA kernels (ggad_csr_row_sum_f64 / ggad_csr_transpose / ggad_csr_scale_add_identity); the results are
    bit-identical to the scipy path (tested against the reference's dense matrices).  ``adj`` is the raw adjacency as
    a scipy matrix, or a CSRGraph already on the device whose fp32 values hold the stored weights exactly.

    pow(deg, -1/2) is the one step whose last bit differs between libm and CUDA, so for integer-valued degrees (every
    dataset of the reference) D comes from a host table indexed by the degree (max_deg + 1 numpy evaluations); for
    non-integer degrees it is evaluated on the device and the fp32 values can differ from scipy's in the last bit."""
    import ctypes as C
    if isinstance(adj, CSRGraph):
        a = adj
    else:
        import scipy.sparse as sp
        m = sp.csr_matrix(adj)
        m.sum_duplicates()
        m.sort_indices()
        a = CSRGraph.from_arrays(m.indptr, m.indices, m.data.astype(np.float32), m.shape[0], m.shape[1], device=device)
    assert a.n_rows == a.n_cols, "adjacency must be square"
    n, dev, h = a.n_rows, a.device, lib()
    with torch.cuda.device(dev):
        st = stream_ptr(dev)
        deg = torch.empty(n, dtype=torch.float64, device=dev)
        check(h.ggad_csr_row_sum_f64(ptr(a.rowptr), ptr(a.val), n, ptr(deg), st))
        deg_i = deg.round()
        if bool((deg_i == deg).all()) and float(deg.max()) < (1 << 24):
            with np.errstate(divide="ignore"):
                table = np.power(np.arange(int(deg.max().item()) + 1, dtype=np.float64), -0.5)   # utils.py:50
            table[np.isinf(table)] = 0.0                                                         # utils.py:51
            dis = torch.from_numpy(table).to(dev)[deg_i.long()]
        else:
            dis = torch.pow(deg, -0.5)
            dis[torch.isinf(dis)] = 0.0
        at = a.T                                         # device transpose (exact); rows keep sorted columns
        sym = bool(torch.equal(a.rowptr, at.rowptr) and torch.equal(a.col, at.col)
                   and (a.val is None or torch.equal(a.val, at.val)))

        def plus_identity(g, scale):
            rp = torch.empty(n + 1, dtype=torch.int64, device=dev)
            nnz = C.c_int64(0)
            check(h.ggad_csr_add_identity_rowptr(ptr(g.rowptr), ptr(g.col), n, ptr(rp), C.addressof(nnz), st))
            col = torch.empty(nnz.value, dtype=torch.int32, device=dev)
            val = torch.empty(nnz.value, dtype=torch.float32, device=dev)
            check(h.ggad_csr_scale_add_identity(ptr(g.rowptr), ptr(g.col), ptr(g.val), ptr(scale), n, ptr(rp), ptr(col),
                                                ptr(val), st))
            return CSRGraph(rp, col, val, n, n, symmetric_pattern=sym)

        g_hat = plus_identity(at, dis)                   # (A D)^T D + I  = D A^T D + I
        g_r = plus_identity(a, None)                     # A + I
    return g_hat, g_r
