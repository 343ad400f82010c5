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
