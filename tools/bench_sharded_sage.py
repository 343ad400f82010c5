#!/usr/bin/env python
"""BASELINE.json config 5: two-layer mean-SAGE mini-batches on the synthetic power-law graph (R-MAT, 6.25 M nodes /
125 M edges per GPU; at 8 GPUs = 50 M nodes / 1 B edges, d = 64), feature table SHARDED by node range with one NCCL
all-reduce of partial accumulators per layer (ggad_b200.sharded) -- next to the replicated-table data-parallel
variant (every rank holds the whole table and graph and draws its own seeds; only the three small parameter
gradients are all-reduced).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded_sage.py [--mode both]

A step = forward (hop blocks, gathers, projections, loss) + backward + gradient sync + Adam for one super-batch of
--seeds seeds.  One JSON line: seeds/s and block edges/s of each mode (max over ranks, CUDA events).
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
    ap.add_argument("--nodes-per-gpu", type=int, default=6_250_000)
    ap.add_argument("--edges-per-gpu", type=int, default=125_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--h", type=int, default=64)
    ap.add_argument("--seeds", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--mode", default="both", choices=["sharded", "replicated", "both"])
    ap.add_argument("--balanced", action="store_true",
                    help="sharded mode: column ranges balanced by adjacency entries (+ 4 per table row) instead of equal node "
                         "ranges.  Measured SLOWER at N=2 (19.2 vs 9.5 ms/step, profiles/r02h_sharded_sage_n2_balanced.json): "
                         "the hot rank's frontier blocks grow with its share of the hub columns; kept as an option")
    args = ap.parse_args()
    import torch.distributed as dist
    from ggad_b200 import _lib, sharded, synth
    from ggad_b200.graph import CSRGraph, DeviceAdjacency
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_local, m_local, d, h = args.nodes_per_gpu, args.edges_per_gpu, args.d, args.h
    n_glob = n_local * world
    # column (= source node) ranges of the sharded mode: balanced by the number of adjacency entries, not by node
    # count -- on a power-law graph the first node range holds most of the edges (R-MAT: 44 % of the sources of the
    # 8-shard graph fall into shard 0), and the slowest rank sets the step time.  Exact integer histogram, all-reduced.
    lo, hi = rank * n_local, (rank + 1) * n_local
    if world > 1 and args.balanced:
        from ggad_b200 import dist as gdist
        mine = synth.rmat_shard(n_local, m_local, world, rank, seed=0, device=dev, mean=False)
        cnt = torch.empty(n_glob, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().ggad_col_histogram(_lib.ptr(mine.col), mine.nnz, _lib.ptr(cnt), n_glob, _lib.stream_ptr(dev)))
        dist.all_reduce(cnt)
        cum = np.zeros(n_glob + 1, np.int64)
        np.cumsum(cnt.cpu().numpy(), out=cum[1:])
        lo, hi = gdist.nnz_balanced_ranges(cum, world, row_cost=4.0)[rank]      # a table row weighs like 4 entries
        del mine, cnt, cum
    out = {"workload": "C5 two-layer mean-SAGE mini-batch", "n_gpus": world, "global_nodes": n_glob,
           "global_edges": m_local * world, "d": d, "h": h, "seeds_per_step": args.seeds}

    def params():
        torch.manual_seed(0)
        ws = [(torch.randn(h, d) * 0.1).to(dev).requires_grad_(True), (torch.randn(h, h) * 0.1).to(dev).requires_grad_(True),
              (torch.randn(2, h) * 0.1).to(dev).requires_grad_(True)]
        return ws, torch.optim.Adam(ws, lr=1e-3)

    def timed(model, opt, seed_fn, sync):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        edges = 0
        for i in range(args.warm + args.iters):
            if i == args.warm:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                ev[0].record()
            seeds, labels = seed_fn(i)
            opt.zero_grad(set_to_none=True)
            loss = model.loss(seeds, labels)
            loss.backward()
            sync(model)
            opt.step()
            if i >= args.warm:
                edges += model.stats["hop1_edges"] + model.stats["hop2_edges"]
        ev[1].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / args.iters, float(edges) / args.iters], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t)
            t[0] = tm[0]
        return float(t[0]), float(t[1]), float(loss), dict(model.stats)

    if args.mode in ("sharded", "both"):
        t0 = time.perf_counter()
        # A[:, lo:hi]: regenerate every destination shard's edges, keep sources in this rank's range, transpose
        g = synth.rmat_transposed_shard(n_local, m_local, world, 0, lo, hi, device=dev)      # rows = local sources
        gt = g.T                                                                             # rows = all destinations
        adj = DeviceAdjacency(gt.rowptr, gt.col, n_glob)
        del g
        x_local = torch.randn(hi - lo, d, device=dev)
        ws, opt = params()
        model = sharded.ShardedTwoLayerSage(sharded.DeviceBackend(adj), x_local, lo, hi, *ws)
        build_s = time.perf_counter() - t0
        common = np.random.default_rng(11)

        def seed_fn(i):                        # the SAME super-batch on every rank
            s = torch.from_numpy(common.integers(0, n_glob, args.seeds))
            return s, (s % 2)
        ms, edges, loss, st = timed(model, opt, seed_fn, lambda m: m.sync_grads())
        out["sharded"] = {"ms_per_step": ms, "seeds_per_s": args.seeds / ms * 1e3, "block_edges_per_step_all_ranks": edges,
                          "block_edges_per_s": edges / ms * 1e3, "frontier_U1": st["u1"], "wire_bytes_per_rank_per_step": 4 * st["wire_floats"],
                          "table_bytes_per_rank": (hi - lo) * d * 4, "adjacency_entries_per_rank": int(adj.col.numel()),
                          "node_range": [int(lo), int(hi)], "loss": loss, "build_s": round(build_s, 1)}
        del model, adj, gt, x_local
        torch.cuda.empty_cache()

    if args.mode in ("replicated", "both"):
        t0 = time.perf_counter()
        parts = [synth.rmat_shard(n_local, m_local, world, s, seed=0, device=dev, mean=False) for s in range(world)]
        offs = np.cumsum([0] + [p.nnz for p in parts])
        rowptr = torch.cat([parts[0].rowptr] + [p.rowptr[1:] + int(offs[i + 1]) for i, p in enumerate(parts[1:])])
        col = torch.cat([p.col for p in parts])
        del parts
        adj = DeviceAdjacency(rowptr, col, n_glob)
        x_full = torch.randn(n_glob, d, device=dev)
        ws, opt = params()
        model = sharded.ShardedTwoLayerSage(sharded.DeviceBackend(adj), x_full, 0, n_glob, *ws)
        model.world = 1                        # no data-path collective: every rank owns everything
        build_s = time.perf_counter() - t0
        own = np.random.default_rng(100 + rank)

        def seed_fn(i):                        # every rank draws its OWN seeds
            s = torch.from_numpy(own.integers(0, n_glob, args.seeds))
            return s, (s % 2)

        def sync(m):
            if world > 1:
                flat = torch.cat([w.grad.reshape(-1) for w in ws])
                dist.all_reduce(flat)
                flat /= world
                o = 0
                for w in ws:
                    w.grad.copy_(flat[o:o + w.numel()].reshape(w.shape))
                    o += w.numel()
        ms, edges, loss, st = timed(model, opt, seed_fn, sync)
        out["replicated"] = {"ms_per_step": ms, "seeds_per_s": world * args.seeds / ms * 1e3, "block_edges_per_step_all_ranks": edges,
                             "block_edges_per_s": edges / ms * 1e3, "frontier_U1": st["u1"], "table_bytes_per_rank": n_glob * d * 4,
                             "adjacency_entries_per_rank": int(adj.col.numel()), "loss": loss, "build_s": round(build_s, 1)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
