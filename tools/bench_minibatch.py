#!/usr/bin/env python
"""Program B (mini-batch GGAD on a DGraph-shaped graph): one training batch = GCN.loss + backward + Adam
(src/model_handler.py:330-364) with the drop-in modules, frontier built on the device.

    python tools/bench_minibatch.py [--nodes 3700550 --edges 36552754 --batch 150 --seeds 50] [--cpu-nodes 300000]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_minibatch.py   (data parallel)

With N ranks the graph, the feature table and the parameters are replicated, every rank draws its own seed batches
and the parameter gradients are averaged with one all-reduce per batch (train.DataParallelMiniBatch): weak scaling,
`batches_per_s` is the aggregate over all ranks.

Also times the CPU restatement of the reference's batch (oracle, sparse; the reference's own dense-mask batch
took 1.52 s on a 300 k-node proxy in the survey) on a smaller graph of the same generator.  One JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=3_700_550)
    ap.add_argument("--edges", type=int, default=36_552_754, help="directed R-MAT candidates (symmetrised afterwards)")
    ap.add_argument("--d", type=int, default=17)
    ap.add_argument("--h", type=int, default=64)
    ap.add_argument("--batch", type=int, default=150)
    ap.add_argument("--seeds", type=int, default=50)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--warm", type=int, default=10, help="untimed batches (frontier sizes vary; the memory pools settle)")
    ap.add_argument("--cpu-nodes", type=int, default=300_000)
    ap.add_argument("--cpu-iters", type=int, default=3)
    ap.add_argument("--reserve-edges", type=int, default=-1,
                    help="hint for DeviceAdjacency.reserve_edges: grow the memory pools ONCE for hop blocks of up to this "
                         "many edges (a new maximum later costs a cudaMalloc, ~1.5 s when the memory is peer-mapped on an "
                         "8-GPU box).  -1 = twice the largest block of the warm-up batches, 0 = no hint")
    ap.add_argument("--graphed", action="store_true",
                    help="train.GraphedMiniBatchStep: dense tail + backward + Adam replayed as one CUDA graph per batch")
    ap.add_argument("--no-prefetch", action="store_true", help="build every batch's hop blocks inside the step (no input pipeline)")
    ap.add_argument("--diag", action="store_true", help="also time the rank-local part of every batch (adds a sync)")
    ap.add_argument("--phases", action="store_true",
                    help="wrap the phases of a batch (hop blocks, gathers, dense tail, backward, optimiser) with "
                         "synchronising timers and report their mean ms (perturbs the total; diagnosis only)")
    args = ap.parse_args()
    from ggad_b200 import _lib, graphsage as gs, synth
    from ggad_b200.train import DataParallelMiniBatch
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    adj = synth.rmat_adjacency(args.nodes, args.edges, seed=72, device=dev)
    adj.reserve_edges = max(0, args.reserve_edges)
    n, d, h = args.nodes, args.d, args.h
    rng = np.random.default_rng(72)
    x = rng.random((n, d), dtype=np.float32)
    x = x / (x.sum(1, keepdims=True) + 0.01)                      # src/utils.py:79 normalisation
    feats = torch.nn.Embedding(n, d)
    feats.weight = torch.nn.Parameter(torch.from_numpy(x), requires_grad=False)
    feats = feats.to(dev)
    agg = gs.GCNAggregator(feats, cuda=True)
    torch.manual_seed(72)                                         # identical initial parameters on every rank
    enc = gs.GCNEncoder(feats, d, h, adj, agg, gcn=True, cuda=True)
    model = gs.GCN(2, enc).to(dev)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.007)
    dp = DataParallelMiniBatch(model, opt)
    graphed = None
    if args.graphed:
        from ggad_b200.train import GraphedMiniBatchStep
        graphed = GraphedMiniBatchStep(model, lr=1e-3, weight_decay=0.007, batch_rows=args.batch + args.seeds)
    rng = np.random.default_rng(72 + 1000 * rank)                 # ... and different seed batches
    deg = (adj.rowptr[1:] - adj.rowptr[:-1]).cpu().numpy()
    cand = np.flatnonzero(deg > 0)
    B = args.batch + args.seeds
    labels = torch.cat([torch.zeros(args.batch, dtype=torch.long), torch.ones(args.seeds, dtype=torch.long)])

    prefetch = None if (args.no_prefetch or args.phases) else gs.BlockPrefetcher(agg, adj)
    upcoming = [rng.choice(cand, B, replace=False).tolist()]
    if prefetch:
        prefetch.submit(upcoming[0])

    def batch(i):
        nodes = upcoming.pop(0)
        upcoming.append(rng.choice(cand, B, replace=False).tolist())
        if prefetch:
            prefetch.submit(upcoming[0])         # batch i+1's frontier is built while batch i trains
        if graphed is not None:
            return graphed.step(nodes, labels)[0]
        if args.diag:
            opt.zero_grad(set_to_none=True)
            t0 = time.perf_counter()
            out = model.loss(nodes, labels)
            out[0].backward()
            torch.cuda.synchronize()
            local_ms.append((time.perf_counter() - t0) * 1e3)
            dp.allreduce_grads()
            opt.step()
            return out[0]
        return dp.step(nodes, labels)[0]

    local_ms = []
    phase_ms = {}
    if args.phases:
        import functools

        def wrap(mod, name, label):
            fn = getattr(mod, name)

            @functools.wraps(fn)
            def timed(*a, **k):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                phase_ms.setdefault(label, []).append((time.perf_counter() - t0) * 1e3)
                return r
            setattr(mod, name, timed)
        wrap(gs, "_block_for", "hop block (frontier + block CSR)")
        wrap(gs, "_aggregate", "gather-reduce from the table")
        wrap(agg, "forward", "aggregator.forward (2 blocks + 2 gathers)")
        wrap(enc, "forward", "encoder.forward (aggregator + projections + ego)")
        wrap(model, "loss", "model.loss (forward)")
        wrap(torch.Tensor, "backward", "backward")
        wrap(opt, "step", "optimizer.step")

    stats = []
    for i in range(args.iters + args.warm):
        if i == args.warm:
            if args.reserve_edges < 0 and stats:
                # settle the memory pools before the timed region: the next block() call (one dummy node) reserves
                # for 2x the hint, i.e. 4x the largest hop-2 block seen so far
                adj.reserve_edges = 2 * int(max(s_[5] for s_ in stats))
                adj.block(torch.tensor([int(cand[0])], dtype=torch.int32, device=dev), True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t_all = time.perf_counter()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        loss = batch(i)
        lv = float(loss.detach())
        torch.cuda.synchronize()
        hop1, hop2 = agg.last_blocks
        stats.append(((time.perf_counter() - t0) * 1e3, _lib.launch_count() - l0, hop1.n_cols,
                      int(hop1.col_d.numel()), -1 if hasattr(hop2, "val_d") else hop2.n_cols, int(hop2.col_d.numel())))
    wall = torch.tensor([time.perf_counter() - t_all], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    st = np.array(stats[args.warm:], dtype=np.float64)
    ms = float(np.median(st[:, 0]))
    agg_bps = world * args.iters / float(wall.item())
    w0 = model.weight.detach().clone()
    if world > 1:                                                  # replicas must stay bit-identical
        wmin, wmax = w0.clone(), w0.clone()
        dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        assert torch.equal(wmin, wmax), "data-parallel replicas diverged"
    edges = float(np.mean(st[:, 3] + st[:, 5]))
    out = {"workload": "C4 mini-batch GGAD", "nodes": n, "adjacency_entries": int(adj.col.numel()), "d": d, "h": h,
           "batch": B, "n_gpus": world, "ms_per_batch": ms, "batches_per_s": agg_bps, "batches_per_s_per_gpu": agg_bps / world,
           "wall_ms_per_lockstep_batch": float(wall.item()) / args.iters * 1e3, "max_ms_per_batch_rank0": float(st[:, 0].max()), "ggad_launches_per_batch": float(np.mean(st[:, 1])),
           "mean_frontier_U": float(np.mean(st[:, 2])), "mean_hop1_edges": float(np.mean(st[:, 3])),
           "mean_frontier_U2": float(np.mean(st[:, 4])), "mean_hop2_edges": float(np.mean(st[:, 5])),
           "edges_per_s": edges * agg_bps, "loss": lv, "prefetch_next_batch_blocks": prefetch is not None, "cuda_graph_tail": graphed is not None,
           "reference_cpu_s_per_batch_300k_proxy_survey": 1.52}
    if args.phases:
        out["phase_ms_mean"] = {k: round(float(np.mean(v[-3 * args.iters // 4:])) * (len(v) / (args.iters + args.warm)), 3) for k, v in phase_ms.items()}
    if args.diag:
        per = {"rank": rank, "local_ms": [round(float(np.percentile(local_ms[args.warm:], q)), 2) for q in (10, 50, 90, 100)],
               "step_ms": [round(float(np.percentile(st[:, 0], q)), 2) for q in (10, 50, 90, 100)]}
        allp = [None] * world
        if world > 1:
            dist.all_gather_object(allp, per)
        else:
            allp = [per]
        out["diag_p10_p50_p90_max"] = allp
    if args.cpu_nodes > 0 and world == 1:
        import oracle
        na = args.cpu_nodes
        adj_c = synth.rmat_adjacency(na, int(args.edges * na / n), seed=72, device=dev)
        rp, col = adj_c.rowptr.cpu().numpy(), adj_c.col.cpu().numpy()
        t0 = time.perf_counter()
        adj_lists = {i: set(col[rp[i]:rp[i + 1]].tolist()) for i in range(na)}
        t_dict = time.perf_counter() - t0
        xc = torch.from_numpy(x[:na].copy())
        p = {"enc.weight": enc.weight.detach().cpu().clone().requires_grad_(True),
             "enc.fc.weight": enc.fc.weight.detach().cpu().clone().requires_grad_(True),
             "weight": model.weight.detach().cpu().clone().requires_grad_(True)}
        degc = np.diff(rp)
        candc = np.flatnonzero(degc > 0)
        tt = []
        for _ in range(args.cpu_iters):
            nodes = rng.choice(candc, B, replace=False).tolist()
            t0 = time.perf_counter()
            total = oracle.gcn_minibatch_loss(p, nodes, labels, adj_lists, xc)[0]
            total.backward()
            tt.append(time.perf_counter() - t0)
        out.update({"cpu_oracle_s_per_batch": float(np.median(tt)), "cpu_graph_nodes": na, "cpu_threads": torch.get_num_threads(),
                    "cpu_dict_build_s": t_dict})
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
