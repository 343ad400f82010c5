"""Training-step drivers that keep the whole step on the device.

``GraphedFullBatchStep`` captures one full-batch GGAD epoch of the reference's loop (run.py:145-213: forward,
BCE + local-affinity margin + reconstruction losses, backward, Adam) into a CUDA graph.  The Photo-sized
configuration moves ~3 us worth of bytes per layer, so the eager epoch is bound by the ~150 kernel launches
of the dense tail; replaying a graph removes that.  The only per-epoch input is the Gaussian noise tensor
(model.py:143), copied into a static buffer before each replay -- results are identical to the eager step.
"""
from __future__ import annotations

import types
from typing import Optional, Sequence

import torch

from ._lib import check, lib, ptr, stream_ptr
from .graph import CSRGraph
from .losses import ggad_loss
from .model import Model, as_graph


class GraphedFullBatchStep:
    def __init__(self, model: Model, features: torch.Tensor, adj, raw_adj, normal_idx: Sequence[int],
                 abnormal_idx: Sequence[int], args, lr: float = 1e-3, weight_decay: float = 0.0,
                 negsamp_ratio: float = 1.0, warmup: int = 3, use_graph: bool = True):
        self.model, self.args = model, args
        dev = features.device
        self.x = features if features.dim() == 3 else features.unsqueeze(0)
        self.adj, self.raw = as_graph(adj, dev), as_graph(raw_adj, dev)
        self.normal, self.abnormal = list(normal_idx), list(abnormal_idx)
        self.negsamp_ratio = negsamp_ratio
        h = model.fc4.in_features
        self.noise = torch.zeros(1, len(self.abnormal), h, device=dev)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=use_graph)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                state = [p.detach().clone() for p in model.parameters()]
                for _ in range(max(1, warmup)):                 # builds plans / transposes / caches, warms the allocator
                    self._eager()
                with torch.no_grad():                           # warm-up must not change the model
                    for p, s in zip(model.parameters(), state):
                        p.copy_(s)
                self.opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=True)
                self._eager()                                   # initialise Adam state outside the capture ...
                with torch.no_grad():
                    for p, s in zip(model.parameters(), state):
                        p.copy_(s)
                for st in self.opt.state.values():              # ... and reset it
                    st["step"].zero_()
                    st["exp_avg"].zero_()
                    st["exp_avg_sq"].zero_()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self.out = self._eager()

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        emb, comb, logits, emb_con, emb_abn = self.model(self.x, self.adj, self.abnormal, self.normal, True, self.args,
                                                         noise=self.noise)
        loss, margin, bce, rec, aff_n, aff_a = ggad_loss(emb, logits, emb_con, emb_abn, self.raw, self.normal,
                                                         self.abnormal, negsamp_ratio=self.negsamp_ratio)
        loss.backward()
        self.opt.step()
        return loss.detach(), margin.detach(), bce.detach(), rec.detach()

    def step(self, noise: Optional[torch.Tensor] = None):
        """One epoch.  ``noise`` ([1,|S|,h] or [|S|,h]); drawn like model.py:143 when omitted.
        Returns (loss, margin, bce, rec) device tensors (static buffers when graphed)."""
        if noise is None:
            noise = torch.randn(self.noise.size()) * self.args.var + self.args.mean
        self.noise.copy_(noise.reshape(self.noise.shape), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            return self.out
        return self._eager()


class GraphedMiniBatchStep:
    """Program B's training batch (src/model_handler.py:330-364) with the dense tail, the backward pass and Adam in
    ONE CUDA-graph launch.

    What varies from batch to batch is integer work -- the frontier U, the hop blocks, their sizes -- and it stays
    eager (device kernels, prefetchable with graphsage.BlockPrefetcher).  Everything after the two table gathers has
    static shapes once the |U|-sized tensors are padded to a capacity: the gathers write straight into static buffers
    (rows beyond |U| are zero and are referenced by no block column, so they contribute exact zeros to every sum and
    gradient), the hop-1 ego-mean operator and its transpose are copied into fixed-size CSR buffers, and one replay
    runs ~150 small kernels (projections, ego mean, outlier generation, the three losses, their backward, Adam) that
    cost 3.3 ms of host time when issued eagerly.  Capacities grow by doubling (one re-capture).  Results equal
    ``model.loss(...).backward(); opt.step()`` (tested)."""

    def __init__(self, model, lr: float = 1e-3, weight_decay: float = 0.0, batch_rows: int = 200, u_cap: int = 1 << 16,
                 e_cap: int = 1 << 18, warmup: int = 2, group=None):
        import torch.distributed as dist
        from . import graphsage as gs
        self.model, self.gs = model, gs
        self.enc, self.agg = model.enc, model.enc.aggregator
        self.dev = gs._device_of(self.enc.features)
        self.lr, self.wd, self.b, self.warmup = lr, weight_decay, int(batch_rows), warmup
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dist, self.group = dist, group
        # with one rank Adam is part of the graph; data parallel: see self.flat below
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay, capturable=True)
        self.graph = None
        self.graph_b = None
        # data parallel: graph A = forward + backward + "flatten the gradients into self.flat"; eager NCCL all-reduce of
        # that one buffer; graph B = average + Adam reading the gradients through views of it -- three launches per batch
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=self.dev) if self.world > 1 else None
        self._alloc(u_cap, e_cap)

    def _alloc(self, u_cap, e_cap):
        dev, b = self.dev, self.b
        table = self.gs._feature_table(self.enc.features)
        dp = table.shape[1]
        self.u_cap, self.e_cap = int(u_cap), int(e_cap)
        self.to_feats = torch.zeros(b, dp, device=dev)
        self.to_feats_neigh = torch.zeros(self.u_cap, dp, device=dev)
        self.lab = torch.zeros(b, dtype=torch.int64, device=dev)
        z64 = lambda n: torch.zeros(n, dtype=torch.int64, device=dev)
        z32 = lambda n: torch.zeros(n, dtype=torch.int32, device=dev)
        # ego-mean operator M = mask / rdeg over the hop-1 block [B, |U|] and its transpose [u_cap, B] (per-edge values)
        self.m_rowptr, self.m_col, self.m_rs = z64(b + 1), z32(self.e_cap), torch.zeros(b, device=dev)
        self.t_rowptr, self.t_col, self.t_val = z64(self.u_cap + 1), z32(self.e_cap), torch.zeros(self.e_cap, device=dev)
        # both run the merge-path tiled kernel (a hub among the batch nodes is one long row) on a plan PADDED to the
        # capacity: the captured launch has a fixed tile count, the plan arrays are refreshed per batch (_load)
        g = CSRGraph(self.m_rowptr, self.m_col, None, b, self.u_cap, row_scale=self.m_rs, use_plan=True)
        gt = CSRGraph(self.t_rowptr, self.t_col, self.t_val, self.u_cap, b, use_plan=True)
        for gg in (g, gt):
            nt = int(lib().ggad_plan_num_tiles(gg.n_rows, self.e_cap))
            gg._plan = (torch.zeros(nt + 1, dtype=torch.int32, device=dev), torch.zeros(nt + 1, dtype=torch.int64, device=dev), nt)
        g._T, gt._T = gt, g
        self.g, self.gt = g, gt
        outer = self

        class _Mask:                       # what loss_from_aggregates calls .mm on / takes the operator from
            def __init__(self):
                self.g = g

            def mm(self, x):
                from . import ops
                return ops.spmm(g, x.contiguous())
        self.mask = _Mask()
        self.graph, self.graph_b, self.out = None, None, None

    def _tail(self):
        for p in self.params:
            p.grad = None
        out = self.model.loss_from_aggregates(self.to_feats[:, :self.enc.feat_dim], self.to_feats_neigh[:, :self.enc.feat_dim],
                                              self.mask, self.lab)
        out[0].backward()
        if self.world == 1:
            self.opt.step()
        else:
            o = 0
            for p in self.params:
                n = p.numel()
                if p.grad is None:
                    self.flat[o:o + n].zero_()
                else:
                    self.flat[o:o + n].copy_(p.grad.reshape(-1))
                o += n
        return tuple(t.detach() for t in out)

    def _views(self):
        """Point every parameter's .grad at its slice of the flat buffer (what graph B's Adam reads)."""
        o = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[o:o + n].view_as(p)
            o += n

    def _apply(self):
        self.flat.div_(self.world)
        self.opt.step()

    def _capture(self):
        dev = self.dev
        state = [p.detach().clone() for p in self.params]
        # a re-capture (capacity growth in the middle of training) must not disturb the optimiser: snapshot its state
        keys = ("step", "exp_avg", "exp_avg_sq")
        snap = {id(p): {k: self.opt.state[p][k].clone() for k in keys} for p in self.params if self.opt.state.get(p)}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, self.warmup)):          # warms the allocator, creates the Adam state tensors
                self._tail()
                if self.world > 1:
                    self._views()
                    self._apply()
            with torch.no_grad():
                for p, s_ in zip(self.params, state):
                    p.copy_(s_)
            for p in self.params:
                st = self.opt.state[p]
                for k in keys:
                    if id(p) in snap:
                        st[k].copy_(snap[id(p)][k])
                    else:
                        st[k].zero_()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._tail()
        if self.world > 1:
            self._views()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b):
                self._apply()

    def _load(self, nodes, labels):
        """Eager, per batch: hop blocks (or the prefetched ones), the two gathers from the feature table into the static
        buffers, the ego-mean operator and its transpose into the fixed-size CSR buffers."""
        from . import ops
        gs, dev = self.gs, self.dev
        ready = self.agg.prefetcher.take(nodes) if getattr(self.agg, "prefetcher", None) is not None else None
        hop1 = ready[0] if ready else gs._block_for(nodes, None, self.enc.adj_lists, True, dev)
        hop2 = ready[1] if ready else gs._hop2_block(hop1.frontier_d, self.enc.adj_lists, self.enc.features, dev)
        self.agg.last_blocks = (hop1, hop2)
        assert hop1.n_rows == self.b, f"GraphedMiniBatchStep was built for {self.b}-node batches, got {hop1.n_rows}"
        e1 = int(hop1.col_d.numel())
        if hop1.n_cols > self.u_cap or e1 > self.e_cap:                     # grow once, re-capture
            self._alloc(max(self.u_cap, 2 * hop1.n_cols), max(self.e_cap, 2 * e1))
        table = gs._feature_table(self.enc.features)
        u = hop1.n_cols
        gs._aggregate(hop1, "sym", self.enc.features, dev, out=self.to_feats)
        self.to_feats_neigh[u:].zero_()
        gs._aggregate(hop2, "sym", self.enc.features, dev, out=self.to_feats_neigh[:u])
        gm = hop1.graph("mean")
        gt = gm.T                                                            # device transpose; 1/rdeg folded into values
        self.m_rowptr.copy_(gm.rowptr)
        self.m_col[:e1].copy_(gm.col)
        self.m_rs.copy_(gm.row_scale)
        self.t_rowptr[:u + 1].copy_(gt.rowptr)
        self.t_rowptr[u + 1:].fill_(e1)                                      # padding rows are empty
        self.t_col[:e1].copy_(gt.col)
        self.t_val[:e1].copy_(gt.val)
        with torch.cuda.device(dev):
            for gg in (self.g, self.gt):
                tr, te, nt = gg._plan
                check(lib().ggad_plan_build_padded(ptr(gg.rowptr), gg.n_rows, e1, nt, ptr(tr), ptr(te), stream_ptr(dev)))
        if not labels.is_cuda and bool(((labels != 0) & (labels != 1)).any()):
            raise RuntimeError("GraphedMiniBatchStep: labels must be 0 / 1 (the static-shape loss; use GCN.loss_reference_path otherwise)")
        self.lab.copy_(labels.reshape(-1), non_blocking=True)

    def step(self, nodes, labels):
        """One batch; returns (total, cls, margin, rec) as static device tensors."""
        self._load(nodes, labels)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        if self.world > 1:
            self.dist.all_reduce(self.flat, group=self.group)
            self.graph_b.replay()
        return self.out


class DataParallelMiniBatch:
    """Data-parallel driver of program B's training batch (src/model_handler.py:330-364) for N GPUs.

    The reference has no multi-GPU path.  On 180 GB parts the DGraph-sized adjacency (0.6 GB) and feature table
    (0.25 GB) are simply replicated; every rank draws its OWN seed batch, runs the drop-in ``GCN.loss`` + backward
    locally (device-side frontier, gather-reduce kernels), and the parameter gradients -- three small tensors --
    are averaged with ONE all-reduce of a flat buffer before Adam.  That is the scheme that scales linearly:
    there is no data-path exchange at all.  Equivalent to one process minimising the mean of the per-rank batch
    losses (tested against exactly that)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, group=None):
        import torch.distributed as dist
        self.model, self.opt, self.group = model, optimizer, group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._flat: Optional[torch.Tensor] = None

    def allreduce_grads(self) -> None:
        """Average the gradients over the ranks: one collective on a flat fp32 buffer (a parameter that got no
        gradient on this rank contributes zeros)."""
        if self.world == 1:
            return
        n = sum(p.numel() for p in self.params)
        if self._flat is None or self._flat.numel() != n or self._flat.device != self.params[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                self._flat[off:off + k].zero_()
            else:
                self._flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        self.dist.all_reduce(self._flat, group=self.group)
        self._flat.div_(self.world)
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                p.grad = self._flat[off:off + k].reshape(p.shape).clone()
            else:
                p.grad.copy_(self._flat[off:off + k].reshape(p.shape))
            off += k

    def step(self, nodes, labels):
        """One batch on this rank's seeds; returns the rank-local (total, cls, margin, rec) losses."""
        self.opt.zero_grad(set_to_none=True)
        out = self.model.loss(nodes, labels)
        out[0].backward()
        self.allreduce_grads()
        self.opt.step()
        return out
