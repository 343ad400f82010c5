"""Drop-in replacement for the reference's ``src/graphsage.py`` (program B, mini-batch GGAD on DGraph).

Same class names, constructor arguments, forward signatures / return tuples and parameter names
(``enc.weight``, ``enc.fc.weight``, ``weight``) as /root/reference/src/graphsage.py, so
``from graphsage import *`` in src/model_handler.py can point here.  What changes underneath:

  * the Python ``set`` unions and the dense 0/1 masks ([B,|U|] and [|U|,|U2|], 1.3 GB per batch on a
    DGraph-sized graph) become a batch-local block CSR with exact integer row / column degrees
    (graph.batch_block);
  * ``mask.mm(embed_matrix)`` becomes the CSR gather-reduce kernel reading the feature table directly
    through an index map (ops.spmm with xmap), with the sym-norm / mean weights applied as row and
    column scales.

The CPU branch of the reference is followed (no ``+ features(nodes)`` residual, src/graphsage.py:325-327).
Frontiers are in sorted-id order where the reference has Python-set order; every consumer is
permutation-invariant.  Tensors must be on a CUDA device.
"""
from __future__ import annotations

import random
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from . import ops
from .graph import AdjListCSR, CSRGraph


# ------------------------------------------------------------------------------------------
# helpers: feature table, batch blocks
# ------------------------------------------------------------------------------------------
def _device_of(features, fallback=None) -> torch.device:
    if isinstance(features, nn.Embedding):
        return features.weight.device
    return fallback or torch.device("cuda")


_table_cache = {}

# GCN.loss: run everything after the two projections as the fused tail kernels (csrc/tail.cu) instead of ~100 small torch
# kernels; GGAD_TORCH_TAIL=1 keeps the torch formulation (the two are parity-tested against each other)
import os as _os
FUSED_TAIL = not _os.environ.get("GGAD_TORCH_TAIL")


def _feature_table(features) -> Optional[torch.Tensor]:
    """The frozen feature table as a 16-byte-row CUDA matrix, or None when ``features`` is a callable."""
    if not isinstance(features, nn.Embedding):
        return None
    w = features.weight
    if not w.is_cuda:
        raise RuntimeError("ggad_b200: move the feature table to a CUDA device (features.cuda()); no CPU fallback")
    hit = _table_cache.get("t")
    if hit is None or hit[0] is not w or hit[1] != (w._version, w.data_ptr(), tuple(w.shape)):
        hit = (w, (w._version, w.data_ptr(), tuple(w.shape)), ops.pad_cols(w.detach()))
        _table_cache["t"] = hit           # one table at a time; the entry keeps `w` alive, so no id recycling
    return hit[2]


class _Block:
    """One aggregation hop of a mini-batch: frontier ids + block CSR on the device, exact integer degrees.
    Built either from host arrays (explicit neighbor sets) or entirely on the device (DeviceAdjacency.block)."""

    def __init__(self, n_rows, n_cols, rowptr_d, col_d, frontier_d, cdeg_i, device):
        self.n_rows, self.n_cols, self.device = int(n_rows), int(n_cols), device
        self.rowptr_d, self.col_d, self.frontier_d = rowptr_d, col_d, frontier_d
        self.rdeg_i = rowptr_d[1:] - rowptr_d[:-1]                 # int64, exact
        self.cdeg_i = cdeg_i                                       # int32, exact
        self.rdeg_d = self.rdeg_i.to(torch.float32)
        self.cdeg_d = cdeg_i.to(torch.float32)

    @classmethod
    def from_host(cls, rowptr: np.ndarray, cols_global: np.ndarray, device) -> "_Block":
        frontier, col_local = np.unique(cols_global, return_inverse=True)
        cdeg = np.bincount(col_local, minlength=len(frontier)).astype(np.int32)
        return cls(len(rowptr) - 1, len(frontier),
                   torch.from_numpy(rowptr.astype(np.int64)).to(device, non_blocking=True),
                   torch.from_numpy(col_local.astype(np.int32)).to(device, non_blocking=True),
                   torch.from_numpy(frontier.astype(np.int32)).to(device, non_blocking=True),
                   torch.from_numpy(cdeg).to(device, non_blocking=True), device)

    @classmethod
    def from_device(cls, adj_dev, nodes_d: torch.Tensor, add_self: bool) -> "_Block":
        b = adj_dev.block(nodes_d, add_self)
        return cls(b["n_rows"], b["n_cols"], b["rowptr"], b["col"], b["frontier"], b["cdeg"], adj_dev.device)

    # host views (tests / API parity); the training path never touches them
    @property
    def frontier(self) -> np.ndarray:
        return self.frontier_d.cpu().numpy().astype(np.int64)

    @property
    def rdeg(self) -> np.ndarray:
        return self.rdeg_i.cpu().numpy()

    @property
    def cdeg(self) -> np.ndarray:
        return self.cdeg_i.cpu().numpy().astype(np.int64)

    def graph(self, mode: str) -> CSRGraph:
        """'sym': 1/sqrt(rdeg) * 1/sqrt(cdeg) (src/graphsage.py:314-318); 'mean': 1/rdeg (:316-317, :92-93).
        Degree-0 rows give 0 * inf = NaN exactly like the reference's 0/0."""
        if mode == "sym":
            rs, cs = 1.0 / self.rdeg_d.sqrt(), 1.0 / self.cdeg_d.sqrt()
        else:
            rs, cs = 1.0 / self.rdeg_d, None
        return CSRGraph(self.rowptr_d, self.col_d, None, self.n_rows, self.n_cols, row_scale=rs, col_scale=cs)


def _rows_from_sets(sets: Sequence[Iterable[int]], nodes: Optional[Sequence[int]], add_self: bool):
    """list[set] (+ optional self union) -> (rowptr, cols) with per-row sorted unique ids."""
    rows_cols = []
    for i, s in enumerate(sets):
        a = np.fromiter((int(t) for t in s), dtype=np.int64, count=len(s))
        if add_self:
            a = np.append(a, int(nodes[i]))
        rows_cols.append(np.unique(a))
    rowptr = np.zeros(len(rows_cols) + 1, dtype=np.int64)
    np.cumsum([len(a) for a in rows_cols], out=rowptr[1:])
    cols = np.concatenate(rows_cols) if rows_cols else np.zeros(0, np.int64)
    return rowptr, cols


class _LazyNeighs(list):
    """What our encoders hand to the aggregators instead of ``[adj_lists[int(n)] for n in nodes]``:
    behaves like that list if someone iterates it, but lets the aggregator slice the cached host CSR."""

    def __init__(self, adj_lists, nodes):
        super().__init__()
        self.adj_lists, self.nodes = adj_lists, [int(n) for n in nodes]

    def materialise(self):
        return [self.adj_lists[n] for n in self.nodes]


def _block_for(nodes, to_neighs, adj_lists, add_self, device) -> _Block:
    """Block of one hop.  Explicit neighbor sets (a caller-supplied ``to_neighs`` list) are honoured on the
    host; otherwise the frontier is built on the device from the cached adjacency CSR."""
    if isinstance(to_neighs, _LazyNeighs) or to_neighs is None:
        adj_dev = adj_lists if hasattr(adj_lists, "block") else AdjListCSR.get(adj_lists).device(device)
        if isinstance(nodes, torch.Tensor):
            nodes_d = nodes
        else:
            nodes_d = torch.as_tensor([int(n) for n in nodes], dtype=torch.int32).to(device, non_blocking=True)
        return _Block.from_device(adj_dev, nodes_d, add_self)
    nodes = [int(n) for n in nodes]
    rowptr, cols = _rows_from_sets(to_neighs, nodes, add_self)
    return _Block.from_host(rowptr, cols, device)


class _DirectBlock:
    """A hop block that is only gathered from the feature table (hop 2 of GCNAggregator): global column ids and the
    batch-local 1/sqrt(cdeg) as per-edge values (DeviceAdjacency.block_direct) -- sym-norm aggregation without the
    frontier list.  ``n_cols`` / ``frontier`` are derived on demand (statistics and tests only)."""

    def __init__(self, b, device):
        self.n_rows, self.device = int(b["n_rows"]), device
        self.rowptr_d, self.col_d, self.val_d, self.n_table = b["rowptr"], b["col"], b["val"], int(b["n_cols"])
        self.rdeg_i = self.rowptr_d[1:] - self.rowptr_d[:-1]
        self.rdeg_d = self.rdeg_i.to(torch.float32)

    @property
    def frontier_d(self):
        return torch.unique(self.col_d)

    @property
    def frontier(self):
        return self.frontier_d.cpu().numpy().astype(np.int64)

    @property
    def n_cols(self):
        return int(self.frontier_d.numel())

    def graph(self, mode: str) -> CSRGraph:
        assert mode == "sym"
        return CSRGraph(self.rowptr_d, self.col_d, self.val_d, self.n_rows, self.n_table, row_scale=1.0 / self.rdeg_d.sqrt())

    def tensors(self):
        return (self.rowptr_d, self.col_d, self.val_d, self.rdeg_i, self.rdeg_d)


def _hop2_block(frontier_d: torch.Tensor, adj_lists, features, device):
    """Hop-2 block of a training batch: the direct form when the features are a table on this device."""
    adj_dev = adj_lists if hasattr(adj_lists, "block") else AdjListCSR.get(adj_lists).device(device)
    if isinstance(features, nn.Embedding) and hasattr(adj_dev, "block_direct"):
        return _DirectBlock(adj_dev.block_direct(frontier_d, False), device)
    return _block_for(frontier_d, None, adj_lists, False, device)


def _aggregate(block, mode: str, features, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(weights of ``mode``) @ features[frontier]  -> [rows, d] (``out``: padded-width destination, returned as is)."""
    g = block.graph(mode)
    if isinstance(block, _DirectBlock):
        table = _feature_table(features)
        if out is not None:
            ops.gather_reduce(g, table, y_out=out)
            return out
        return ops.spmm(g, table)[:, :features.weight.shape[1]]
    if out is not None:
        ops.gather_reduce(g, _feature_table(features), xmap=block.frontier_d, y_out=out)
        return out
    table = _feature_table(features)
    if table is not None:
        d = features.weight.shape[1]
        return ops.spmm(g, table, xmap=block.frontier_d)[:, :d]
    embed = features(block.frontier_d.long())                                   # callable (stacked encoders)
    return ops.spmm(g, embed)


class BlockMask:
    """The row-mean mask the reference returns as a dense [B,|U|] tensor (src/graphsage.py:316-317,360);
    only ``.mm`` is consumed by GCNEncoder (:421)."""

    def __init__(self, block: _Block):
        self.block = block
        self._g = block.graph("mean")
        self.shape = (block.n_rows, block.n_cols)

    def mm(self, x: torch.Tensor) -> torch.Tensor:
        return ops.spmm(self._g, x.contiguous())

    def to_dense(self) -> torch.Tensor:
        b = self.block
        m = torch.zeros(self.shape, device=b.device)
        rows = torch.repeat_interleave(torch.arange(b.n_rows, device=b.device), b.rdeg_i)
        m[rows, b.col_d.long()] = (1.0 / b.rdeg_d)[rows]
        return m


# ------------------------------------------------------------------------------------------
# vanilla GraphSAGE (src/graphsage.py:19-154)
# ------------------------------------------------------------------------------------------
class GraphSage(nn.Module):
    def __init__(self, num_classes, enc):
        super(GraphSage, self).__init__()
        self.enc = enc
        self.xent = nn.CrossEntropyLoss()
        self.weight = nn.Parameter(torch.FloatTensor(num_classes, enc.embed_dim))
        init.xavier_uniform_(self.weight)

    def forward(self, nodes):
        return ops.linear(self.enc(nodes).t(), self.weight)

    def to_prob(self, nodes):
        return torch.sigmoid(self.forward(nodes))

    def loss(self, nodes, labels):
        scores = self.forward(nodes)
        return self.xent(scores, labels.squeeze().to(scores.device))


class MeanAggregator(nn.Module):
    """Mean of (optionally sampled) neighbor features (src/graphsage.py:46-99)."""

    def __init__(self, features, cuda=False, gcn=False):
        super(MeanAggregator, self).__init__()
        self.features = features
        self.cuda = cuda
        self.gcn = gcn
        self.adj_lists = None            # set by Encoder so the cached host CSR can be sliced

    def forward(self, nodes, to_neighs, num_sample=10):
        device = _device_of(self.features)
        if num_sample is not None:
            sets = to_neighs.materialise() if isinstance(to_neighs, _LazyNeighs) else to_neighs
            # host sampling like the reference (:74-80); sorted() because random.sample rejects sets on py>=3.11
            to_neighs = [set(random.sample(sorted(s), num_sample)) if len(s) >= num_sample else s for s in sets]
        block = _block_for(nodes, to_neighs, self.adj_lists, self.gcn, device)
        return _aggregate(block, "mean", self.features, device)


class Encoder(nn.Module):
    """ReLU(W . cat(self, mean-of-neighbors)^T)  (src/graphsage.py:102-154)."""

    def __init__(self, features, feature_dim, embed_dim, adj_lists, aggregator, num_sample=10, base_model=None,
                 gcn=False, cuda=False, feature_transform=False):
        super(Encoder, self).__init__()
        self.features = features
        self.feat_dim = feature_dim
        self.adj_lists = adj_lists
        self.aggregator = aggregator
        self.num_sample = num_sample
        if base_model != None:
            self.base_model = base_model
        self.gcn = gcn
        self.embed_dim = embed_dim
        self.cuda = cuda
        self.aggregator.cuda = cuda
        self.aggregator.adj_lists = adj_lists
        self.weight = nn.Parameter(torch.FloatTensor(embed_dim, self.feat_dim if self.gcn else 2 * self.feat_dim))
        init.xavier_uniform_(self.weight)

    def forward(self, nodes):
        neigh_feats = self.aggregator.forward(nodes, _LazyNeighs(self.adj_lists, nodes), self.num_sample)
        if not self.gcn:
            index = torch.as_tensor(nodes, dtype=torch.long).to(neigh_feats.device)
            combined = torch.cat((self.features(index), neigh_feats), dim=1)
        else:
            combined = neigh_feats
        return ops.linear(combined, self.weight, relu=True).t()          # ReLU(W . combined^T)  (:153)


# ------------------------------------------------------------------------------------------
# GGAD mini-batch model (src/graphsage.py:157-454)
# ------------------------------------------------------------------------------------------
class GCNAggregator(nn.Module):
    """Two-hop symmetric-normalised aggregation with batch-local degrees (src/graphsage.py:275-360)."""

    def __init__(self, features, cuda=False, gcn=False):
        super(GCNAggregator, self).__init__()
        self.features = features
        self.cuda = cuda
        self.gcn = gcn

    def forward(self, nodes, to_neighs, adj_list, train_flag):
        device = _device_of(self.features)
        ready = self.prefetcher.take(nodes) if (getattr(self, "prefetcher", None) is not None and train_flag == True) else None
        hop1 = ready[0] if ready else _block_for(nodes, to_neighs, adj_list, True, device)    # N(b) U {b}   (:305)
        to_feats = _aggregate(hop1, "sym", self.features, device)
        to_feats_neigh = None
        if train_flag == True:
            hop2 = ready[1] if ready else _hop2_block(hop1.frontier_d, adj_list, self.features, device)   # no self union (:335)
            to_feats_neigh = _aggregate(hop2, "sym", self.features, device)
            self.last_blocks = (hop1, hop2)
        else:
            self.last_blocks = (hop1, None)
        return to_feats, to_feats_neigh, BlockMask(hop1)


class BlockPrefetcher:
    """Input pipeline of program B's training loop: the two hop blocks of the NEXT batch (frontier union, block CSR,
    batch-local degrees -- integer work that ends in two device->host size reads per hop) are built by a helper thread
    on its own CUDA stream while the current batch runs its gathers, dense tail, backward and optimiser step.
    The reference builds them with Python set unions inside ``GCNAggregator.forward`` (src/graphsage.py:305-311,
    335-341); results are identical (same kernels), only the schedule changes.

        pf = BlockPrefetcher(aggregator, adj_lists);  pf.submit(batch_0)
        for i: pf.submit(batch_{i+1});  model.loss(batch_i, labels_i) ...     # forward picks the prepared blocks up
    """

    def __init__(self, aggregator: "GCNAggregator", adj_lists, switch_interval: float = 2e-4):
        import sys
        import threading
        if switch_interval and sys.getswitchinterval() > switch_interval:
            sys.setswitchinterval(switch_interval)     # the helper needs the GIL for microseconds at a time; 5 ms slices starve it
        self.agg, self.adj_lists = aggregator, adj_lists
        self.device = _device_of(aggregator.features)
        self.stream = torch.cuda.Stream(device=self.device)
        self.pending = {}
        self.lock = threading.Lock()
        aggregator.prefetcher = self

    def _build(self, nodes, slot):
        try:
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                hop1 = _block_for(nodes, None, self.adj_lists, True, self.device)
                hop2 = _hop2_block(hop1.frontier_d, self.adj_lists, self.agg.features, self.device)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            slot["result"] = (hop1, hop2, ev)
        except BaseException as e:           # surfaced by take()
            slot["error"] = e

    def submit(self, nodes) -> None:
        import threading
        slot = {}
        slot["thread"] = threading.Thread(target=self._build, args=(nodes, slot), daemon=True)
        with self.lock:
            self.pending[id(nodes)] = (nodes, slot)
        slot["thread"].start()

    def take(self, nodes):
        with self.lock:
            hit = self.pending.pop(id(nodes), None)
        if hit is None or hit[0] is not nodes:
            return None
        slot = hit[1]
        slot["thread"].join()
        if "error" in slot:
            raise slot["error"]
        hop1, hop2, ev = slot["result"]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for b in (hop1, hop2):               # allocated on the helper stream, consumed on this one
            ts = b.tensors() if hasattr(b, "tensors") else (b.rowptr_d, b.col_d, b.frontier_d, b.cdeg_i, b.rdeg_i, b.rdeg_d, b.cdeg_d)
            for t in ts:
                t.record_stream(cur)
        return hop1, hop2


class GCNEncoder(nn.Module):
    """src/graphsage.py:363-454.  Parameters: weight [h,d], fc.weight [h,h]."""

    def __init__(self, features, feature_dim, embed_dim, adj_lists, aggregator, num_sample=10, base_model=None,
                 gcn=False, cuda=False, feature_transform=False):
        super(GCNEncoder, self).__init__()
        self.features = features
        self.feat_dim = feature_dim
        self.adj_lists = adj_lists
        self.aggregator = aggregator
        self.num_sample = num_sample
        if base_model != None:
            self.base_model = base_model
        self.gcn = gcn
        self.embed_dim = embed_dim
        self.cuda = cuda
        self.aggregator.cuda = cuda
        self.weight = nn.Parameter(torch.FloatTensor(embed_dim, self.feat_dim))
        init.xavier_uniform_(self.weight)
        self.fc = nn.Linear(embed_dim, embed_dim, bias=False)

    def forward(self, nodes, label, train_flag):
        neigh_feats, neigh_feats_expand, mask = self.aggregator.forward(nodes, _LazyNeighs(self.adj_lists, nodes),
                                                                        self.adj_lists, train_flag)
        combined = ops.linear(neigh_feats, self.weight, relu=True).t()           # aggregate, then project (:412)
        to_feats_neigh = None
        anomaly_feat = None
        anomaly_feat_new = None
        combined_all = combined
        if train_flag == True:
            label = label.to(combined.device)
            emb_u = ops.linear(neigh_feats_expand, self.weight, relu=True)       # hop-1 frontier embeddings (:419), [|U|, h]
            to_feats_neigh = mask.mm(emb_u)                                      # ego-neighbor mean (:421)
            is_ab, is_norm = label == 1, label == 0
            anomaly_feat = combined[:, is_ab]
            anomaly_feat_new = ops.linear(to_feats_neigh[is_ab], self.fc.weight, relu=True)   # outlier generation (:428-430)
            combined_all = torch.cat((combined[:, is_norm], anomaly_feat_new.t()), 1)   # label-0 columns first (:450)
            anomaly_feat_new = anomaly_feat_new.t()
        return combined_all, to_feats_neigh, anomaly_feat, anomaly_feat_new


class GCN(nn.Module):
    """src/graphsage.py:157-272: scores, local-affinity margin, reconstruction, total loss."""

    def __init__(self, num_classes, enc):
        super(GCN, self).__init__()
        self.enc = enc
        self.xent = nn.BCEWithLogitsLoss(reduction='none', pos_weight=torch.tensor([1]))
        self.weight = nn.Parameter(torch.FloatTensor(1, enc.embed_dim))
        init.xavier_uniform_(self.weight)

    def forward(self, nodes, label, train_flag):
        embeds, to_feats_neigh, anomaly_feat, anomaly_feat_new = self.enc(nodes, label, train_flag)
        scores = ops.linear(embeds.t(), self.weight)                             # (weight . embeds)^T  (:174)
        return scores, to_feats_neigh, embeds, anomaly_feat, anomaly_feat_new

    def to_prob(self, nodes, label):
        return torch.sigmoid(self.forward(nodes, label, train_flag=False)[0])

    def to_prob_reconstruction(self, nodes, label):
        return self.forward(nodes, label, train_flag=False)[0]

    def recon2(self, anomaly_feat, anomaly_feat_new):
        return torch.mean(torch.sqrt(torch.sum(torch.pow(anomaly_feat - anomaly_feat_new, 2), 0)))

    def normalize(self, emb):
        inv = torch.pow(torch.norm(emb, dim=-1, keepdim=True), -1)
        inv = torch.where(torch.isinf(inv), torch.zeros_like(inv), inv)
        return emb * inv

    def affinity(self, combined_all, labels, to_feats_neigh):
        labels = labels.to(combined_all.device)
        aff = torch.cosine_similarity(combined_all, to_feats_neigh.t(), dim=0)   # (:234)
        aff_normal = torch.mean(aff[torch.argwhere(labels == 0)], 0)
        aff_abnormal = torch.mean(aff[torch.argwhere(labels == 1)], 0)
        return (1 - (aff_normal - aff_abnormal)).clamp_min(min=0)                # margin 1 (:236-240)

    def loss(self, nodes, labels):
        """total, cls, margin, rec of src/graphsage.py:244-258 -- same numbers as composing forward / affinity / recon2
        (``loss_reference_path`` below, kept for the parity test), but written with label MASKS and one stable sort
        instead of boolean indexing: every tensor has a static shape, so the step issues no device->host reads after
        the frontier is built (boolean indexing synchronises five times per batch)."""
        enc = self.enc
        lab = labels.detach()
        if not lab.is_cuda:
            if bool(((lab != 0) & (lab != 1)).any()):           # CPU labels: free to inspect
                return self.loss_reference_path(nodes, labels)
        neigh_feats, neigh_feats_expand, mask = enc.aggregator.forward(nodes, _LazyNeighs(enc.adj_lists, nodes),
                                                                       enc.adj_lists, True)
        return self.loss_from_aggregates(neigh_feats, neigh_feats_expand, mask, lab.to(neigh_feats.device).reshape(-1))

    def loss_from_aggregates(self, neigh_feats, neigh_feats_expand, mask, lab):
        """The dense tail of a training batch from the two aggregates ([B,d] and [|U|,d], rows beyond |U| may be zero
        padding), the ego-mean mask (anything with ``.mm``) and the device label vector: every shape is static, which
        is what train.GraphedMiniBatchStep captures into a CUDA graph."""
        enc = self.enc
        g_mean = getattr(mask, "_g", None) or getattr(mask, "g", None)
        if FUSED_TAIL and g_mean is not None and enc.embed_dim <= 256 and enc.embed_dim % 4 == 0:
            # projections on the GEMM kernel, everything after them in two fused launches (ops.minibatch_tail)
            combined = ops.linear(neigh_feats, enc.weight, relu=True)            # [B,h]  (:412)
            emb_u = ops.linear(neigh_feats_expand, enc.weight, relu=True)        # [|U|,h] (:419)
            total, cls, margin, rec = ops.minibatch_tail(combined, emb_u, enc.fc.weight, self.weight, lab, g_mean)
            return total, cls, margin, rec
        is_ab = lab == 1
        m1 = is_ab.to(torch.float32)
        m0 = (lab == 0).to(torch.float32)
        combined = ops.linear(neigh_feats, enc.weight, relu=True)                # [B,h]  (:412)
        emb_u = ops.linear(neigh_feats_expand, enc.weight, relu=True)            # [|U|,h] (:419)
        ego = mask.mm(emb_u)                                                     # [B,h]  (:421)
        afn_all = ops.linear(ego, enc.fc.weight, relu=True)                      # ReLU(fc(ego)) for every row; rows with label 1 are consumed (:430)
        # combined_all (:450): label-0 columns in batch order, then the generated outliers -> stable sort of the labels
        order = torch.sort(is_ab.to(torch.int8), stable=True).indices
        rows_all = torch.where(is_ab[order].unsqueeze(1), afn_all[order], combined[order])   # combined_all^T, [B,h]
        scores = ops.linear(rows_all, self.weight)                               # (:174)
        loss_cls = torch.mean(self.xent(scores.squeeze(), lab.to(torch.float32)))   # labels stay in batch order (:246)
        aff = torch.cosine_similarity(rows_all, ego, dim=1)                      # column p against ego row p (:234)
        aff_normal = ((aff * m0).sum() / m0.sum()).reshape(1)
        aff_abnormal = ((aff * m1).sum() / m1.sum()).reshape(1)
        loss_constraint = (1 - (aff_normal - aff_abnormal)).clamp_min(min=0)     # (:236-240)
        sq = torch.sum(torch.pow(combined - afn_all, 2), 1)
        dist_ = torch.sqrt(torch.where(is_ab, sq, torch.ones_like(sq)))          # rows without label 1 never reach sqrt'(0)
        loss_rec = (dist_ * m1).sum() / m1.sum()                                 # recon2 (:197-198)
        return 1 * loss_cls + 1 * loss_constraint + 0.1 * loss_rec, loss_cls, loss_constraint, loss_rec

    def loss_reference_path(self, nodes, labels):
        scores, to_feats_neigh, embeds, anomaly_feat, anomaly_feat_new = self.forward(nodes, labels, train_flag=True)
        target = labels.detach().to(scores.device, dtype=torch.float32)
        loss_cls = torch.mean(self.xent(scores.squeeze(), target))
        loss_constraint = self.affinity(embeds, labels, to_feats_neigh)
        loss_rec = self.recon2(anomaly_feat, anomaly_feat_new)
        return 1 * loss_cls + 1 * loss_constraint + 0.1 * loss_rec, loss_cls, loss_constraint, loss_rec
