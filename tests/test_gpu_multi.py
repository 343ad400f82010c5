"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the node-range sharded layer pass over NCCL equals
the single-GPU pass on the concatenated graph (rtol 1e-5: only the fp32 summation association differs),
integer partition arrays are identical on every rank."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ggad_b200 import _lib, dist as gdist, ops, synth
    from ggad_b200.graph import CSRGraph
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n_local, m_local, d, seed = 40000, 700000, 64, 5
    n_glob = n_local * world
    fwd = synth.rmat_shard(n_local, m_local, world, rank, seed=seed, device=dev, mean=True)
    fr = [(g * n_local, (g + 1) * n_local) for g in range(world)]
    cnt = torch.empty(n_glob, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ggad_col_histogram(_lib.ptr(fwd.col), fwd.nnz, _lib.ptr(cnt), n_glob, _lib.stream_ptr(dev)))
    dist.all_reduce(cnt)
    rowptr_t = np.zeros(n_glob + 1, np.int64)
    np.cumsum(cnt.cpu().numpy(), out=rowptr_t[1:])
    br = gdist.nnz_balanced_ranges(rowptr_t, world, row_cost=1)
    rs_all = torch.empty(n_glob, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(rs_all, fwd.row_scale)
    lo, hi = br[rank]
    bwd = synth.rmat_transposed_shard(n_local, m_local, world, seed, lo, hi, device=dev, col_scale=rs_all).fold_col_scale()
    gen = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(n_glob, d, device=dev, generator=gen)
    lp = gdist.ShardedLayerPass(fwd, bwd, fr, br, rank, lambda g, t: ops.gather_reduce(g, t)["y"])
    y = lp.forward(x)
    dx = lp.backward(y)
    ok, msg = True, ""
    # fused exchange: P2P stores from the gather epilogue (and NVSwitch multicast when available) must give
    # bit-identical replicas to the NCCL all-gather of the same kernel's output
    rep = gdist.PeerReplica(n_glob, d, fr, rank, dev)
    rep.barrier(0)
    ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_peers=rep.peer_row_ptrs)
    rep.barrier(1)
    fused_ok = torch.equal(rep.buf, y)
    mc_ok = True
    if rep.multicast_ptr:
        rep.barrier(0)
        rep.buf.zero_()
        rep.barrier(1)
        ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_multicast=rep.multicast_row_ptr)
        rep.barrier(0)
        mc_ok = torch.equal(rep.buf, y)
    # halo exchange: only rows some peer's backward shard gathers cross NVLink.  Needed rows must be bit-identical
    # to the all-gather, rows nobody asked for must be left untouched (poisoned first), and the backward on the
    # halo replica must be bit-identical to the backward on the full replica
    need = gdist.halo_need_mask(bwd.col, fr, rank)
    rep.barrier(0)
    rep.buf.fill_(float("nan"))
    rep.barrier(1)
    ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_peers=rep.peer_row_ptrs, peer_need=need)
    rep.barrier(0)
    mine = torch.zeros(n_glob, dtype=torch.bool, device=dev)
    mine[bwd.col.long()] = True
    mine[fr[rank][0]:fr[rank][1]] = True
    halo_ok = torch.equal(rep.buf[mine], y[mine]) and bool(torch.isnan(rep.buf[~mine]).all())
    dx_halo = ops.gather_reduce(bwd, rep.buf)["y"]
    halo_ok = halo_ok and torch.equal(dx_halo, dx[lo:hi])
    frac = float((need != 0).float().mean())
    rep.barrier(1)
    ok = fused_ok and mc_ok and halo_ok
    msg = f"fused={fused_ok} multicast={'n/a' if not rep.multicast_ptr else mc_ok} halo={halo_ok} (rows sent {frac:.2f}) "
    if rank == 0:
        # single-GPU reference on the concatenated graph, same kernels
        parts = [synth.rmat_shard(n_local, m_local, world, s, seed=seed, device=dev, mean=True) for s in range(world)]
        rowptr = torch.cat([parts[0].rowptr] + [p.rowptr[1:] + sum(q.nnz for q in parts[:i + 1])
                                                for i, p in enumerate(parts[1:])])
        full = CSRGraph(rowptr, torch.cat([p.col for p in parts]), None, n_glob, n_glob,
                        row_scale=torch.cat([p.row_scale for p in parts]))
        y_ref = ops.gather_reduce(full, x)["y"]
        dx_ref = ops.gather_reduce(full.T, y_ref)["y"]
        ok = ok and torch.allclose(y, y_ref, rtol=1e-5, atol=1e-6) and torch.allclose(dx, dx_ref, rtol=1e-5, atol=1e-5)
        msg += f"max|dy|={float((y - y_ref).abs().max()):.3e} max|ddx|={float((dx - dx_ref).abs().max()):.3e}"
        # exact integer check: the transposed shards tile the global transpose
        assert int(bwd.nnz) > 0
    tot = torch.tensor([bwd.nnz], device=dev)
    dist.all_reduce(tot)
    ok = ok and int(tot.item()) == m_local * world
    q.put((rank, bool(ok), msg, br))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_layer_pass_nccl():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res
    assert all(r[3] == res[0][3] for r in res)
