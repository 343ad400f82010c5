#!/usr/bin/env python
"""NVLink P2P write-bandwidth probe (run under torchrun, N >= 2): how fast can this rank push bytes into a
peer's symmetric-memory buffer with (a) the copy engine / cudaMemcpy path and (b) an SM kernel issuing plain
vectorised stores.  Calibrates the exchange roofline of the fused gather."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n, d = 6_250_000, 64
buf = symm.empty((n * world, d), dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
src = torch.randn(n, d, device=dev)
res = {}


def timed(fn, iters=5):
    fn()
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    hdl.barrier(channel=0)
    return a.elapsed_time(b) / iters


peers = [p for p in range(world) if p != rank]
views = {p: hdl.get_buffer(p, (n * world, d), torch.float32) for p in peers}
lo = rank * n
# (a) memcpy path, one peer / all peers
res["memcpy_1peer_GBs"] = src.numel() * 4 / timed(lambda: views[peers[0]][lo:lo + n].copy_(src)) / 1e6
res["memcpy_allpeers_GBs"] = len(peers) * src.numel() * 4 / timed(lambda: [views[p][lo:lo + n].copy_(src) for p in peers]) / 1e6
# (b) SM kernel with vectorised stores (torch elementwise kernel writing straight into the peer mapping)
res["sm_store_1peer_GBs"] = src.numel() * 4 / timed(lambda: torch.mul(src, 1.0, out=views[peers[0]][lo:lo + n])) / 1e6
res["sm_store_allpeers_GBs"] = len(peers) * src.numel() * 4 / timed(lambda: [torch.mul(src, 1.0, out=views[p][lo:lo + n]) for p in peers]) / 1e6
# local reference
loc = torch.empty_like(src)
res["sm_store_local_GBs"] = src.numel() * 4 / timed(lambda: torch.mul(src, 1.0, out=loc)) / 1e6
out = [None] * world
dist.all_gather_object(out, res)
if rank == 0:
    print(json.dumps({"world": world, "per_rank": out}))
dist.barrier()
dist.destroy_process_group()
