"""torch.autograd.Function wrappers around the C-ABI kernels (K1-K6 of DESIGN.md).

Every op takes CUDA tensors and calls libggad_b200.so on torch's current stream; there is no
CPU path.  Backward passes are hand-derived and call the same gather-reduce kernel on the
transposed CSR (no float atomics anywhere, results are run-to-run deterministic).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GatherDesc, check, lib, ptr, stream_ptr
from .graph import CSRGraph


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def pad_cols(x: torch.Tensor) -> torch.Tensor:
    """Zero-pad the last dim to a multiple of 4 floats (16-byte rows); no copy if already aligned."""
    d = x.shape[-1]
    if d % 4 == 0 and x.is_contiguous() and x.data_ptr() % 16 == 0:
        return x
    out = x.new_zeros(*x.shape[:-1], _pad4(d))
    out[..., :d] = x
    return out


def gather_reduce(g: CSRGraph, x: torch.Tensor, *, xmap: Optional[torch.Tensor] = None,
                  col_scale: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
                  bias: Optional[torch.Tensor] = None, prelu_slope: Optional[torch.Tensor] = None, relu: bool = False,
                  want_y: bool = True, want_z: bool = False, want_sumsq: bool = False,
                  dot_mat: Optional[torch.Tensor] = None, dot_rows: Optional[torch.Tensor] = None,
                  dot_scale: Optional[torch.Tensor] = None, use_graph_scales: bool = True,
                  y_out: Optional[torch.Tensor] = None, y_peers=None, y_multicast: Optional[int] = None,
                  peer_need: Optional[torch.Tensor] = None, chase: bool = False, chase_ctas: int = 0,
                  mc_min_peers: int = 0):
    """Raw (non-autograd) call of ggad_gather_reduce.  ``x`` is [n_x_rows, d] fp32 CUDA with d % 4 == 0.
    Returns dict(y=, z=, sumsq=, dot=) with the requested outputs."""
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.dim() == 2
    x = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) else x.contiguous()
    d = x.shape[1]
    if d % 4 != 0:
        raise RuntimeError(f"ggad_b200: feature width {d} must be a multiple of 4 (use ops.pad_cols)")
    dev = x.device
    n = g.n_rows
    if use_graph_scales:
        row_scale = g.row_scale if row_scale is None else row_scale
        col_scale = g.col_scale if col_scale is None else col_scale
    # shape contract (the device code trusts it: a short operand would be a silent out-of-bounds gather where the
    # reference's bmm / spmm raises a shape error)
    if xmap is not None:
        if xmap.numel() < g.n_cols or xmap.dtype != torch.int32 or not xmap.is_cuda:
            raise RuntimeError(f"ggad_b200: xmap must be an int32 CUDA vector with >= n_cols={g.n_cols} entries")
    elif x.shape[0] < g.n_cols:
        raise RuntimeError(f"ggad_b200: operand has {x.shape[0]} rows but the graph has {g.n_cols} columns")
    for nm, t, need in (("row_scale", row_scale, n), ("col_scale", col_scale, g.n_cols), ("dot_rows", dot_rows, n),
                        ("dot_scale", dot_scale, n), ("bias", bias, d)):
        if t is not None and (t.numel() < need or not t.is_cuda):
            raise RuntimeError(f"ggad_b200: {nm} needs >= {need} CUDA elements, got {t.numel()} on {t.device}")
    if y_out is not None:
        assert y_out.shape == (n, d) and y_out.dtype == torch.float32 and y_out.stride(0) == d and y_out.is_cuda
    y = (y_out if y_out is not None else torch.empty(n, d, dtype=torch.float32, device=dev)) if want_y else None
    z = torch.empty(n, d, dtype=torch.float32, device=dev) if want_z else None
    ss = torch.empty(n, dtype=torch.float32, device=dev) if want_sumsq else None
    dot = torch.empty(n, dtype=torch.float32, device=dev) if dot_mat is not None else None
    desc = GatherDesc()
    desc.rowptr, desc.col, desc.val = ptr(g.rowptr), ptr(g.col), ptr(g.val)
    desc.n_rows, desc.nnz = g.n_rows, g.nnz
    desc.x, desc.ldx = ptr(x), x.stride(0)
    desc.n_x_rows = min(int(x.shape[0]), 2**31 - 1)
    desc.xmap, desc.col_scale, desc.row_scale = ptr(xmap), ptr(col_scale), ptr(row_scale)
    desc.d, desc.relu = d, int(bool(relu))
    desc.bias, desc.prelu_slope = ptr(bias), ptr(prelu_slope)
    desc.y, desc.z, desc.ldy = ptr(y), ptr(z), d
    desc.sumsq = ptr(ss)
    if dot_mat is not None:
        assert dot_mat.stride(1) == 1
        desc.dot_mat, desc.lddot = ptr(dot_mat), dot_mat.stride(0)
        desc.dot_rows, desc.dot_scale, desc.dot_out = ptr(dot_rows), ptr(dot_scale), ptr(dot)
    if y_peers:                       # fused exchange: epilogue also stores into the peers' replicas
        assert len(y_peers) <= 7
        for i, pp in enumerate(y_peers):
            desc.y_peer[i] = int(pp)
        desc.n_peer = len(y_peers)
        if peer_need is not None:     # halo exchange: int32 bit mask per row, bit p = y_peers[p] gathers the row
            assert peer_need.dtype == torch.int32 and peer_need.numel() == n and peer_need.is_cuda
            desc.peer_need = ptr(peer_need)
    if y_multicast and not chase:
        desc.y_multicast = int(y_multicast)
        if y_peers and mc_min_peers > 0:          # hybrid: rows many peers need take the multicast address
            desc.mc_min_peers = int(mc_min_peers)
    plan = g.plan
    if plan is not None:
        ws = g.workspace(d)
        desc.tile_row, desc.tile_edge, desc.n_tiles, desc.ws = ptr(plan[0]), ptr(plan[1]), plan[2], ptr(ws)
    if not (chase and y_peers):
        with torch.cuda.device(dev):
            check(lib().ggad_gather_reduce(desc, stream_ptr(dev)))
        return dict(y=y, z=z, sumsq=ss, dot=dot)
    # ---- chase mode: the gather only flags finished tiles; the exchange runs beside it as ggad_halo_chase on a
    # second, higher-priority stream (enqueued AFTER the gather, so a serialising launch order is still correct)
    if plan is None or y is None:
        raise RuntimeError("ggad_b200: chase exchange needs the merge-path plan and the local y")
    done, epoch = g.chase_flags()
    desc.tile_done, desc.tile_epoch = ptr(done), epoch
    cd = _lib.ChaseDesc()
    cd.y, cd.ldy, cd.d, cd.n_peer = ptr(y), d, d, len(y_peers)
    cd.rowptr, cd.n_rows = ptr(g.rowptr), g.n_rows
    cd.tile_row, cd.tile_edge, cd.n_tiles = ptr(plan[0]), ptr(plan[1]), plan[2]
    cd.tile_done, cd.tile_epoch, cd.n_ctas = ptr(done), epoch, int(chase_ctas)
    cd.peer_need = ptr(peer_need)
    for i, pp in enumerate(y_peers):
        cd.y_peer[i] = int(pp)
    if y_multicast and mc_min_peers > 0:
        cd.y_multicast, cd.mc_min_peers = int(y_multicast), int(mc_min_peers)
    cur = torch.cuda.current_stream(dev)
    side = _chase_stream(dev)
    with torch.cuda.device(dev):
        side.wait_stream(cur)                       # everything the gather waits for, the chase waits for too
        check(lib().ggad_gather_reduce(desc, cur.cuda_stream))
        check(lib().ggad_halo_chase(cd, side.cuda_stream))
        y.record_stream(side)
        cur.wait_stream(side)
    return dict(y=y, z=z, sumsq=ss, dot=dot)


_chase_streams = {}


def _chase_stream(dev) -> torch.cuda.Stream:
    key = torch.device(dev).index
    s = _chase_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev, priority=-1)   # high priority: its few CTAs take the next free SM slots
        _chase_streams[key] = s
    return s


# ------------------------------------------------------------------------------------------
# K6: dense projections (tcgen05 fp32-accurate GEMM / SIMT for tiny or unaligned shapes)
# ------------------------------------------------------------------------------------------
def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D fp32 CUDA view with unit inner stride and a 16-byte row pitch where possible (no copy if it already is)."""
    if t.stride(1) == 1 and t.stride(0) >= t.shape[1] and t.data_ptr() % 4 == 0:
        return t
    return t.contiguous()


def dense_matmul(a: torch.Tensor, b: torch.Tensor, *, trans_a: bool = False, trans_b: bool = False, relu: bool = False,
                 out: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0, path: int = 0) -> torch.Tensor:
    """C = act(alpha * op(a) @ op(b) + beta * C) through ggad_dense_matmul (row-major fp32 CUDA matrices; see
    include/ggad_b200.h).  The result has a leading dimension padded to a multiple of 4 floats (returned as a view),
    so chains of projections stay on the tensor-core path."""
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    a, b = _rowmajor(a), _rowmajor(b)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    kb, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if kb > k and not trans_a and trans_b:
        raise RuntimeError(f"ggad_b200: inner dimensions differ ({k} vs {kb})")
    k = min(k, kb)                      # operands zero-padded along k may differ in their padded extent
    if out is None:
        buf = torch.empty(m, _pad4(n), dtype=torch.float32, device=a.device)
        out = buf[:, :n]
    else:
        assert out.shape == (m, n) and out.stride(1) == 1 and out.is_cuda and out.dtype == torch.float32
    if k == 0 or m == 0 or n == 0:          # nothing to multiply: C = act(beta * C)
        if beta == 0.0:
            out.zero_()
        else:
            out.mul_(beta)
        return torch.relu_(out) if relu else out
    with torch.cuda.device(a.device):
        check(lib().ggad_dense_matmul(int(trans_a), int(trans_b), m, n, k, ptr(a), a.stride(0), ptr(b), b.stride(0),
                                      ptr(out), out.stride(0), float(alpha), float(beta), int(relu), int(path),
                                      stream_ptr(a.device)))
    return out


class _Linear(torch.autograd.Function):
    """y = act(x @ w^T) for a torch nn.Linear weight w [out, in] (no bias) -- model.py:27,156,176-180 and
    src/graphsage.py:412,419,430,174.  Backward: dx = dy @ w, dw = dy^T @ x (both through ggad_dense_matmul)."""

    @staticmethod
    def forward(ctx, x, w, relu: bool):
        k = x.shape[1]
        xp, wp = pad_cols(x), pad_cols(w)          # inner dimension zero-padded to 16-byte rows (745 -> 748): TMA pitch
        y = dense_matmul(xp, wp, trans_b=True, relu=relu)
        ctx.relu, ctx.k = relu, k
        ctx.save_for_backward(xp, wp, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, wp, y = ctx.saved_tensors
        if ctx.relu:
            dy = dy * (y > 0)
        dx = dense_matmul(dy, wp)[:, :ctx.k] if ctx.needs_input_grad[0] else None
        dw = dense_matmul(dy, xp, trans_a=True)[:, :ctx.k] if ctx.needs_input_grad[1] else None
        return dx, dw, None


def linear(x: torch.Tensor, w: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """Drop-in for ``F.linear(x, w)`` (+ optional fused ReLU) on [..., in] inputs."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = _Linear.apply(x2, w, relu)
    return y.reshape(*lead, w.shape[0])


class _MinibatchTail(torch.autograd.Function):
    """The dense tail of a mini-batch GGAD training batch as two fused launches (ggad_minibatch_tail_fwd / _bwd,
    csrc/tail.cu; formulas in include/ggad_b200.h).  Inputs: combined [B,h] (post-ReLU projection), emb_u [|U|,h]
    (post-ReLU hop-1 frontier embeddings), the ego-mean operator (CSRGraph [B,|U|], row_scale 1/rdeg), fc [h,h],
    weight [1,h], labels [B] in {0,1}.  Returns the four loss tensors of src/graphsage.py:244-258; only ``total``
    is differentiable."""

    @staticmethod
    def forward(ctx, combined, emb_u, fc, weight, lab, g_mean):
        dev, (b, h) = combined.device, combined.shape
        ego = gather_reduce(g_mean, emb_u.contiguous())["y"]                     # [B,h] ego-neighbor mean (:421)
        comb = _rowmajor(combined)
        f32 = dict(dtype=torch.float32, device=dev)
        st = dict(rows=torch.empty(b, h, **f32), apre_src=torch.empty(b, h, **f32), apre_own=torch.empty(b, h, **f32),
                  scores=torch.empty(b, **f32), bce=torch.empty(b, **f32), cos=torch.empty(b, **f32),
                  dist=torch.empty(b, **f32), norms=torch.empty(2 * b, **f32),
                  src=torch.empty(b, dtype=torch.int32, device=dev), out=torch.empty(8, **f32))
        d = _lib.TailDesc()
        d.combined, d.ld_combined, d.ego, d.ld_ego = ptr(comb), comb.stride(0), ptr(ego), ego.stride(0)
        fcc, wc, labc = fc.contiguous(), weight.reshape(-1).contiguous(), lab.to(torch.int64).contiguous()
        d.fc, d.weight, d.labels, d.batch, d.h = ptr(fcc), ptr(wc), ptr(labc), b, h
        for k, v in st.items():
            setattr(d, k, ptr(v))
        with torch.cuda.device(dev):
            check(lib().ggad_minibatch_tail_fwd(d, stream_ptr(dev)))
        ctx.g_mean, ctx.st = g_mean, st
        ctx.save_for_backward(comb, ego, fcc, wc, labc, emb_u)
        out = st["out"]
        total, cls, margin, rec = out[0:1].clone(), out[1].clone(), out[2:3].clone(), out[3].clone()
        ctx.mark_non_differentiable(cls, margin, rec)
        return total, cls, margin, rec

    @staticmethod
    def backward(ctx, g_total, *_unused):
        comb, ego, fcc, wc, labc, emb_u = ctx.saved_tensors
        dev, (b, h), st = comb.device, comb.shape, ctx.st
        f32 = dict(dtype=torch.float32, device=dev)
        d_comb, d_apre, d_ego = torch.empty(b, h, **f32), torch.empty(b, h, **f32), torch.empty(b, h, **f32)
        d_scores, d_w = torch.empty(b, **f32), torch.empty(h, **f32)
        d = _lib.TailDesc()
        d.combined, d.ld_combined, d.ego, d.ld_ego = ptr(comb), comb.stride(0), ptr(ego), ego.stride(0)
        d.fc, d.weight, d.labels, d.batch, d.h = ptr(fcc), ptr(wc), ptr(labc), b, h
        for k, v in st.items():
            setattr(d, k, ptr(v))
        gt = g_total.reshape(-1).to(torch.float32).contiguous()
        d.grad_total, d.d_combined, d.ld_d_combined = ptr(gt), ptr(d_comb), h
        d.d_apre, d.d_ego, d.d_scores, d.d_weight = ptr(d_apre), ptr(d_ego), ptr(d_scores), ptr(d_w)
        with torch.cuda.device(dev):
            check(lib().ggad_minibatch_tail_bwd(d, stream_ptr(dev)))
        d_fc = dense_matmul(d_apre, ego, trans_a=True) if ctx.needs_input_grad[2] else None       # d fc = d_apre^T ego
        d_emb = gather_reduce(ctx.g_mean.T, d_ego)["y"] if ctx.needs_input_grad[1] else None       # through the mean operator
        return (d_comb if ctx.needs_input_grad[0] else None), d_emb, d_fc, \
            (d_w.reshape(1, -1) if ctx.needs_input_grad[3] else None), None, None


def minibatch_tail(combined, emb_u, fc, weight, lab, g_mean: CSRGraph):
    """(total [1], cls, margin [1], rec) of a mini-batch GGAD batch from the projected aggregates (see _MinibatchTail)."""
    return _MinibatchTail.apply(combined, emb_u, fc, weight, lab, g_mean)


def halo_push(y: torch.Tensor, y_peers, peer_need: Optional[torch.Tensor] = None) -> None:
    """Store the rows of ``y`` [n, d] (a rank's own block of a replicated matrix) into the peers' replicas over
    NVLink: row r goes to ``y_peers[p]`` (peer-mapped device addresses of the same block) iff bit p of
    ``peer_need[r]`` is set (None = every row to every peer).  Stand-alone form of the gather kernel's
    end-of-tile push, for matrices that no gather launch produced (layer-0 features)."""
    import ctypes as C
    _lib.require_cuda(y, "y")
    assert y.dtype == torch.float32 and y.dim() == 2 and y.stride(1) == 1 and y.shape[1] % 4 == 0
    n_peer = len(y_peers)
    if n_peer == 0 or y.shape[0] == 0:
        return
    if peer_need is not None:
        assert peer_need.dtype == torch.int32 and peer_need.numel() == y.shape[0] and peer_need.is_cuda
    arr = (C.c_void_p * n_peer)(*[int(p) for p in y_peers])
    with torch.cuda.device(y.device):
        check(lib().ggad_halo_push(ptr(y), y.stride(0), y.shape[0], y.shape[1], ptr(peer_need), arr, n_peer,
                                   stream_ptr(y.device)))


# ------------------------------------------------------------------------------------------
class _Spmm(torch.autograd.Function):
    """y = A x  (optionally x gathered through xmap).  dx = A^T dy."""

    @staticmethod
    def forward(ctx, x, g: CSRGraph, xmap):
        ctx.g, ctx.xmap, ctx.n_x = g, xmap, x.shape[0]
        return gather_reduce(g, x, xmap=xmap)["y"]

    @staticmethod
    def backward(ctx, dy):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        g = ctx.g
        if ctx.xmap is not None:
            raise RuntimeError("ggad_b200: gradient w.r.t. a feature table gathered through xmap is not supported "
                               "(the reference freezes it: src/model_handler.py:263-264)")
        dx = gather_reduce(g.T, dy.contiguous())["y"]
        return dx, None, None


def spmm(g: CSRGraph, x: torch.Tensor, xmap: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A @ x on the GPU (replaces torch.spmm / torch.bmm(adj, .) of model.py:29,31 and the
    mask.mm(...) products of src/graphsage.py).  Width is padded to a multiple of 4 internally."""
    d = x.shape[1]
    xp = pad_cols(x)
    y = _Spmm.apply(xp, g, xmap)
    return y if xp.shape[1] == d else y[:, :d]


class _GcnAggregate(torch.autograd.Function):
    """out = PReLU(A x + bias) in one launch (model.py:29-35); saves the pre-activation for backward."""

    @staticmethod
    def forward(ctx, x, g: CSRGraph, bias, slope):
        r = gather_reduce(g, x, bias=bias, prelu_slope=slope, want_z=True)
        ctx.g = g
        ctx.save_for_backward(r["z"], slope)
        ctx.has_bias = bias is not None
        return r["y"]

    @staticmethod
    def backward(ctx, dy):
        z, slope = ctx.saved_tensors
        neg = ~(z > 0)                      # torch's PReLU backward: x > 0 ? g : w * g  (so z == 0 takes the slope)
        dz = torch.where(neg, dy * slope, dy)
        dslope = (dy * z * neg).sum().reshape(slope.shape) if ctx.needs_input_grad[3] else None
        dbias = dz.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = gather_reduce(ctx.g.T, dz)["y"] if ctx.needs_input_grad[0] else None
        return dx, None, dbias, dslope


def gcn_aggregate(g: CSRGraph, x: torch.Tensor, bias: Optional[torch.Tensor], slope: torch.Tensor) -> torch.Tensor:
    d = x.shape[1]
    xp = pad_cols(x)
    bp = None if bias is None else pad_cols(bias.reshape(1, -1)).reshape(-1)
    y = _GcnAggregate.apply(xp, g, bp, slope)
    return y if xp.shape[1] == d else y[:, :d]


class _LocalAffinity(torch.autograd.Function):
    """aff_i = (1/c_j) sum_i' R[i',j] <e^_i', e^_j>, j = subset[i]   (run.py:175-188), only for the rows
    in ``subset``.  Forward = row normalisation + one gather-reduce of e^ over CSR(R^T)[subset] with the dot
    epilogue.  Backward, first term through the transposed row-subset CSR:
        de^_k = sum_j R[k,j] (g_j/c_j) e^_j  +  [k in subset] (g_k/c_k) sum_i' R[i',k] e^_i'
        de_k  = (de^_k - e^_k <e^_k, de^_k>) / |e_k|
    """

    @staticmethod
    def forward(ctx, emb, g_rt_sub: CSRGraph, g_r: CSRGraph, subset, r_inv_sub):
        n, d = emb.shape
        inv = torch.empty(n, dtype=torch.float32, device=emb.device)
        with torch.cuda.device(emb.device):
            check(lib().ggad_row_inv_norm(ptr(emb), emb.stride(0), n, d, ptr(inv), None, stream_ptr(emb.device)))
        # e^ = e / |e| once (N x d elementwise), so the gather runs in the plain per-edge-value mode instead of the
        # general one (a col_scale lookup per edge): aff_j = r_inv_j <sum_i R[i,j] e^_i , e^_j>
        ehat = emb * inv.unsqueeze(1)
        r = gather_reduce(g_rt_sub, ehat, dot_mat=ehat, dot_rows=subset, dot_scale=r_inv_sub, use_graph_scales=False)
        ctx.g_sub = g_rt_sub
        ctx.save_for_backward(emb, inv, ehat, r["y"], subset, r_inv_sub)
        return r["dot"]

    @staticmethod
    def backward(ctx, g_aff):
        emb, inv, ehat, acc, subset, r_inv_sub = ctx.saved_tensors
        n, d = emb.shape
        gamma = g_aff * r_inv_sub                                  # g_j / c_j on the subset
        # first term: sum_{j in subset} R[k,j] gamma_j e^_j  =  (R^T[subset,:])^T @ (gamma * e^[subset]) -- a plain gather
        # over the |subset|/N fraction of the edges (the transposed row-subset CSR, built once and cached), instead of
        # a masked pass over all of R
        xs = ehat[subset.long()] * gamma.unsqueeze(1)
        de = gather_reduce(ctx.g_sub.T, xs, use_graph_scales=False)["y"]
        de.index_add_(0, subset.long(), acc * gamma.unsqueeze(1))  # second term, only on the subset rows
        with torch.cuda.device(emb.device):
            check(lib().ggad_normalize_backward(ptr(emb), emb.stride(0), ptr(inv), ptr(de), de.stride(0), n, d,
                                                stream_ptr(emb.device)))
        return de, None, None, None, None


def local_affinity(emb: torch.Tensor, g_r: CSRGraph, subset: torch.Tensor) -> torch.Tensor:
    """Local-affinity scores for the nodes in ``subset`` (int32 CUDA, unique).  ``g_r`` is R = A + I."""
    d = emb.shape[1]
    e = pad_cols(emb)
    cache = g_r.__dict__.setdefault("_aff_cache", {})
    key = (subset.data_ptr(), subset.numel())
    hit = cache.get(key)
    if hit is None:
        g_rt = g_r.T
        g_sub = g_rt.rows(subset.cpu().numpy())
        # c_j = sum_i R[i,j]: exact column sums (row sums of R^T), fp32 like torch.sum(raw_adj, 0)
        ones = torch.ones(g_r.n_rows, 4, dtype=torch.float32, device=emb.device)
        csum = gather_reduce(g_sub, ones, use_graph_scales=False)["y"][:, 0]
        r_inv = torch.where(csum != 0, 1.0 / csum, torch.zeros_like(csum))
        hit = (g_sub, r_inv.contiguous(), subset)
        cache.clear()
        cache[key] = hit
    g_sub, r_inv, _ = hit
    return _LocalAffinity.apply(e, g_sub, g_r, subset, r_inv)
