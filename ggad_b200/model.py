"""Drop-in replacement for the reference's ``model.py`` (program A, full-batch GGAD).

Same class names, constructor arguments, forward signatures, return tuples and state_dict keys as
/root/reference/model.py, so ``from model import Model`` in run.py can point here unchanged.  The
N x N dense products are replaced by the CSR gather-reduce kernels:

  GCN.forward        model.py:26-35    -> ops.gcn_aggregate (A_hat @ (X W^T) + bias, PReLU: one launch)
  adj[0,S,:] @ emb   model.py:151-155  -> ops.spmm on the row-extracted CSR (ego-neighbor sum)

``adj`` may be a CSRGraph, a scipy matrix, a torch sparse tensor or the dense [1,N,N] tensor run.py
builds (converted once and cached).  Everything must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import CSRGraph

_graph_cache = {}
_index_cache = {}


def _probe(idx):
    """Content key of an index list (the same bytes CSRGraph.rows() keys on), so a list mutated in place can never
    be served stale device indices."""
    import numpy as np
    return hash(np.asarray(idx, dtype=np.int64).tobytes())


def _device_index(idx, device) -> torch.Tensor:
    """LongTensor copy of a Python index list on the device.  Cached per list OBJECT (the cache keeps a
    reference, so an id can never be recycled while its entry is alive) so that the forward pass issues no
    host->device copies -- required for CUDA-graph capture."""
    if isinstance(idx, torch.Tensor):
        return idx.to(device=device, dtype=torch.long)
    key = (id(idx), str(device))
    hit = _index_cache.get(key)
    if hit is None or hit[0] is not idx or hit[2] != _probe(idx):
        t = torch.as_tensor(list(idx), dtype=torch.long).to(device)
        hit = (idx, t, _probe(idx))
        if len(_index_cache) > 16:
            _index_cache.clear()
        _index_cache[key] = hit
    return hit[1]


_noise_stage = {}


def _stage_noise(cpu_noise: torch.Tensor, device) -> torch.Tensor:
    key = (tuple(cpu_noise.shape), str(device))
    slot = _noise_stage.get(key)
    if slot is None:
        slot = [torch.empty(cpu_noise.shape, dtype=cpu_noise.dtype).pin_memory() for _ in range(2)] + [0, None]
        if len(_noise_stage) > 8:
            _noise_stage.clear()
        _noise_stage[key] = slot
    slot[2] ^= 1
    pin = slot[slot[2]]
    if slot[3] is not None:
        slot[3].synchronize()           # the copy issued two calls ago from this buffer has long finished
    pin.copy_(cpu_noise)
    out = pin.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    slot[3] = ev
    return out


def as_graph(adj, device) -> CSRGraph:
    """Convert whatever run.py hands over into a (cached) CSRGraph."""
    if isinstance(adj, CSRGraph):
        return adj
    key = (id(adj), str(device))
    hit = _graph_cache.get(key)
    version = getattr(adj, "_version", None)
    if hit is None or hit[0] is not adj or hit[1] != version:
        hit = (adj, version, CSRGraph.from_any(adj, device))
        if len(_graph_cache) > 8:
            _graph_cache.clear()
        _graph_cache[key] = hit
    return hit[2]


class GCN(nn.Module):
    """model.py:6-35.  Parameters: fc.weight [out,in], bias [out], act.weight [1]."""

    def __init__(self, in_ft, out_ft, act, bias=True):
        super(GCN, self).__init__()
        self.fc = nn.Linear(in_ft, out_ft, bias=False)
        self.act = nn.PReLU() if act == 'prelu' else act
        if bias:
            self.bias = nn.Parameter(torch.FloatTensor(out_ft))
            self.bias.data.fill_(0.0)
        else:
            self.register_parameter('bias', None)
        for m in self.modules():
            self.weights_init(m)

    def weights_init(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight.data)
            if m.bias is not None:
                m.bias.data.fill_(0.0)

    def forward(self, seq, adj, sparse=False):
        squeeze = seq.dim() == 3
        x = seq[0] if squeeze else seq
        g = as_graph(adj, x.device)
        seq_fts = ops.linear(x, self.fc.weight)                # project first (model.py:27), tcgen05 GEMM
        if isinstance(self.act, nn.PReLU) and self.act.weight.numel() == 1:
            out = ops.gcn_aggregate(g, seq_fts, self.bias, self.act.weight)
        else:
            out = ops.spmm(g, seq_fts)
            if self.bias is not None:
                out = out + self.bias
            out = self.act(out)
        return out.unsqueeze(0) if squeeze else out


class _Readout(nn.Module):
    """Placeholder for the reference's four parameter-free readout modules (model.py:38-74).  Model constructs one
    and never calls it (the GGAD path has no graph-level readout); it holds no parameters, so state_dict keys and
    the RNG stream are unaffected."""

    def __init__(self, mode):
        super().__init__()
        self.mode = mode

    def forward(self, seq, query=None):
        if self.mode == 'weighted_sum':
            w = F.softmax(torch.matmul(seq, query.transpose(1, 2)), dim=1)
            return (seq * w.expand(-1, -1, seq.shape[2])).sum(1)
        red = {'avg': torch.mean, 'max': torch.amax, 'min': torch.amin}[self.mode]
        return red(seq, 1)


class Discriminator(nn.Module):
    """Only the constructor matters: Model creates it so that ``disc.f_k.{weight,bias}`` exist in the state_dict and
    the seeded initialisation order matches (model.py:76-90,131).  The reference never calls it on the GGAD path."""

    def __init__(self, n_h, negsamp_round):
        super(Discriminator, self).__init__()
        self.f_k = nn.Bilinear(n_h, n_h, 1)
        torch.nn.init.xavier_uniform_(self.f_k.weight.data)
        if self.f_k.bias is not None:
            self.f_k.bias.data.fill_(0.0)
        self.negsamp_round = negsamp_round


class EncoderModel(nn.Module):
    """The two-layer GCN encoder the reference's other full-batch detectors build from the same ``GCN`` layer
    (``model_ocgnn.py:109-131``: ``Model(n_in, n_h, activation, negsamp_round, readout)``, ``forward(seq1, adj,
    sparse=False) -> h_2``; ``model_AEGIS.py:153-156`` stacks the same layer as encoder / decoder).  Same members in the
    same order, so a reference checkpoint loads with ``strict=True``; both layers run on the gather-reduce path."""

    def __init__(self, n_in, n_h, activation, negsamp_round, readout):
        super(EncoderModel, self).__init__()
        self.read_mode = readout
        self.gcn1 = GCN(n_in, n_h, activation)
        self.gcn2 = GCN(n_h, n_h, activation)
        self.act = nn.ReLU()
        if readout in ('max', 'min', 'avg', 'weighted_sum'):
            self.read = _Readout(readout)
        self.disc = Discriminator(n_h, negsamp_round)

    def forward(self, seq1, adj, sparse=False):
        g = as_graph(adj, seq1.device)
        return self.gcn2(self.gcn1(seq1, g, sparse), g, sparse)


class Model(nn.Module):
    """model.py:108-191.  Unused members (gcn3, fc5, fc6, read, disc) are created in the reference's
    order so seeded initialisation and state_dict keys match."""

    def __init__(self, n_in, n_h, activation, negsamp_round, readout):
        super(Model, self).__init__()
        self.read_mode = readout
        self.gcn1 = GCN(n_in, n_h, activation)
        self.gcn2 = GCN(n_h, n_h, activation)
        self.gcn3 = GCN(n_h, n_h, activation)
        self.fc1 = nn.Linear(n_h, int(n_h / 2), bias=False)
        self.fc2 = nn.Linear(int(n_h / 2), int(n_h / 4), bias=False)
        self.fc3 = nn.Linear(int(n_h / 4), 1, bias=False)
        self.fc4 = nn.Linear(n_h, n_h, bias=False)
        self.fc6 = nn.Linear(n_h, n_h, bias=False)
        self.fc5 = nn.Linear(n_h, n_in, bias=False)
        self.act = nn.ReLU()
        if readout in ('max', 'min', 'avg', 'weighted_sum'):
            self.read = _Readout(readout)
        self.disc = Discriminator(n_h, negsamp_round)

    def _mlp(self, t):
        # f_3 = fc3(ReLU(fc2(ReLU(fc1(t)))))  (model.py:176-180), ReLU fused into the projection epilogue
        return ops.linear(ops.linear(ops.linear(t, self.fc1.weight, relu=True), self.fc2.weight, relu=True), self.fc3.weight)

    def forward(self, seq1, adj, sample_abnormal_idx, normal_idx, train_flag, args, sparse=False, noise=None):
        """Returns (emb, emb_combine, f_3, emb_con, emb_abnormal) exactly like the reference.
        ``noise`` (optional, [1,|S|,h]) overrides the Gaussian draw of model.py:143 for parity tests."""
        g = as_graph(adj, seq1.device)
        h_1 = self.gcn1(seq1, g, sparse)
        emb = self.gcn2(h_1, g, sparse)
        emb_con = None
        emb_combine = None
        s_idx = _device_index(sample_abnormal_idx, emb.device)
        emb_abnormal = emb[:, s_idx, :]
        if noise is None:
            # drawn on the CPU generator with the reference's call (same stream of numbers, model.py:143), staged in a
            # reusable pinned buffer and copied without blocking the host
            noise = _stage_noise(torch.randn(emb_abnormal.size()) * args.var + args.mean, emb.device)
        emb_abnormal = emb_abnormal + noise
        if train_flag:
            ego = ops.spmm(g.rows(sample_abnormal_idx), emb[0])          # rows S of A_hat @ emb
            emb_con = ops.linear(ego, self.fc4.weight, relu=True)              # ReLU(fc4(.)), model.py:155-156
            emb_combine = torch.cat((emb[:, _device_index(normal_idx, emb.device), :], torch.unsqueeze(emb_con, 0)), 1)
            f_3 = self._mlp(emb_combine)
            emb = emb.index_copy(1, s_idx, emb_con.unsqueeze(0))         # the in-place write-back of model.py:182
        else:
            f_3 = self._mlp(emb)
        return emb, emb_combine, f_3, emb_con, emb_abnormal
