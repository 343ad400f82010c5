#!/usr/bin/env python
"""Headline benchmark: edges/s of one CSR SpMM layer pass (forward + backward) on R-MAT power-law graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload S64|C4|custom] [--impl native|reference]

Workload (config.workload): ``S64`` = one C5 shard per GPU -- 6.25 M nodes / 125 M edges / d = 64 per GPU
(at N = 8 this is BASELINE.json's 50 M-node / 1 B-edge graph) -- weak scaling.  A step is
    Y  = A[lo:hi, :] X          (mean aggregation over the rank's destination range), delivered to the ranks
                                whose backward shard gathers it (N>1: NVLink P2P stores fused into the gather
                                epilogue -- only the halo rows by default -- or a plain NCCL all-gather)
    dX[lo':hi'] = A^T[lo':hi', :] dY   (dY = Y, loss = |Y|^2/2; transposed CSR rows, no atomics; stays sharded)
`value` = total edges of all ranks / step time with everything resident in HBM (inputs >> L2, so no flush
is needed); `e2e` = same pass with the features coming from pinned host memory and dX + loss read back,
through the C-ABI host entry point (ggad_spmm_fwd_bwd_host) at N = 1.  `roofline` is for the forward
gather kernel: algorithmic bytes (SURVEY.md 8d B_alg) / CUDA-event time / measured HBM peak.
`--impl reference` times the reference's own CPU path for this op (torch.spmm on a CSR adjacency,
model.py:28-29, + autograd backward) on a bounded sample of the same generator.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nodes per GPU, edges per GPU, width)
    "S64": (6_250_000, 125_000_000, 64),
    "C4": (3_700_550, 73_105_508, 20),          # DGraph-shaped full-graph pass, d=17 padded to 20
    "C3": (39_357, 21_222_543, 300),            # T-Finance-shaped layer-2 pass (L2-resident table)
    "C2": (11_944, 4_398_392, 300),             # Amazon-shaped layer-2 pass
    "tiny": (50_000, 1_000_000, 64),
    # program B's training batch on the DGraph-shaped graph (BASELINE.json config 4): metric = mini-batches/s,
    # replicated graph + table, every rank its own seed batches (tools/bench_minibatch.py does the work)
    "C4mb": (3_700_550, 36_552_754, 17),
}
RMAT = (0.57, 0.19, 0.19)


# ----------------------------------------------------------------------------------------------
# host-side R-MAT (numpy) for the CPU legs -- same distribution as the device generator
# ----------------------------------------------------------------------------------------------
def rmat_csr_numpy(n: int, m: int, seed: int, chunk: int = 1 << 24):
    """R-MAT CSR on the host with the device generator's scheme (per-level quadrant draw, rejection of ids >= n for
    up to 32 rounds, then modulo), vectorised: float32 draws, int32 bit accumulation in 16 M-edge chunks and ONE
    int64 key sort -- the full S64 graph (125 M edges) takes ~1 min instead of the ~4 min of a lexsort."""
    rng = np.random.default_rng(seed)
    bits = max(1, int(np.ceil(np.log2(n))))
    a, b, c = (np.float32(t) for t in RMAT)
    t1, t2, t3 = a, np.float32(a + b), np.float32(a + b + c)
    keys = np.empty(m, dtype=np.int64)

    def draw(k):
        d_ = np.zeros(k, dtype=np.int32)
        s_ = np.zeros(k, dtype=np.int32)
        for _l in range(bits):
            u = rng.random(k, dtype=np.float32)
            q0, q1, q2 = u >= t1, u >= t2, u >= t3          # quadrant = q0 + q1 + q2: 0 a, 1 b, 2 c, 3 d
            np.left_shift(d_, 1, out=d_)
            d_ |= q1                                         # destination bit = quadrant >> 1
            np.left_shift(s_, 1, out=s_)
            s_ |= (q0 & ~q1) | q2                            # source bit = quadrant & 1
        return d_, s_

    for lo in range(0, m, chunk):
        k = min(chunk, m - lo)
        dst, src = draw(k)
        todo = np.flatnonzero((dst >= n) | (src >= n))
        for _ in range(31):
            if len(todo) == 0:
                break
            d2, s2 = draw(len(todo))
            dst[todo], src[todo] = d2, s2
            todo = todo[(d2 >= n) | (s2 >= n)]
        if len(todo):
            dst[todo] %= n
            src[todo] %= n
        keys[lo:lo + k] = (dst.astype(np.int64) << 32) | src
    keys.sort()
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(keys >> 32, minlength=n), out=rowptr[1:])
    return rowptr, (keys & 0xffffffff).astype(np.int32)


def cpu_reference_leg(n: int, m: int, d: int, steps: int, warmup: int, budget_s: float = None):
    """The reference's CPU implementation of the op (torch.spmm on a CSR adjacency, model.py:28-29, + autograd's
    backward) on all host threads.  The operator is built once (run.py builds its adjacency before the epoch loop);
    a step is spmm + backward.  With ``budget_s`` the timed steps stop early once the budget is spent (>= 3 kept)."""
    import oracle
    # all the host threads the process may use (torchrun exports OMP_NUM_THREADS=1 to its workers, which would make
    # the reference arm single-threaded at N > 1)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    t_gen = time.perf_counter()
    rowptr, col = rmat_csr_numpy(n, m, seed=0)
    deg = np.diff(rowptr)
    val = np.repeat(np.where(deg > 0, 1.0 / np.maximum(deg, 1), 0).astype(np.float32), deg)
    a = oracle.torch_spmm_cpu_operator(rowptr, col, val, n)
    del rowptr, col, val
    x = torch.randn(n, d)
    t_gen = time.perf_counter() - t_gen
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.torch_spmm_cpu_step(a, x, backward=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and len(times) >= min(3, steps) and time.perf_counter() - t_begin > budget_s:
            break
    t = float(np.mean(times))
    return dict(value=m / t, unit="edges/s", cores=torch.get_num_threads(), kind="port",
                sample=f"R-MAT N={n} nnz={m} d={d} (same generator as the GPU workload), torch.spmm CSR fwd+bwd, "
                       f"{len(times)} steps of {t:.2f} s on {torch.get_num_threads()} threads / {os.cpu_count()} cpus "
                       f"(graph built once in {t_gen:.0f} s, untimed)"), t, len(times)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region: NVML polled from a thread every
    few ms (the region of a 20-step run is ~100 ms, too short for `nvidia-smi -lms`); falls back to an
    nvidia-smi subprocess if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.proc, self.lines, self.samples, self.th = None, [], [], None
        self.stop_flag = threading.Event()
        self.nvml = None
        self.t_begin, self.t_end = 0.0, float("inf")

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self._physical_index()), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, rs, time.perf_counter()))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.th.join(2)
            nv = self.nvml
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "hw_power_brake_slowdown": nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            inside = [x for x in self.samples if self.t_begin <= x[2] <= self.t_end] or self.samples
            sm = [x[0] for x in inside]
            reasons = sorted(k for k, bit in names.items() if any(x[1] & bit for x in inside))
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.max_mhz, reasons=reasons,
                        samples=len(sm), source="nvml")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
def verify_sampled_rows(g, table: torch.Tensor, out: torch.Tensor, n_samples: int, seed: int):
    """Parity check at the benchmark's own size: recompute ``n_samples`` rows of out = A @ table on the CPU in fp64
    from this rank's rowptr / col (/ val, row_scale) slices -- sample = first, last and highest-degree row + uniform
    draws -- and compare with the kernel's rows.  Bound per element: 1e-4 * |ref| + 1e-5 * sum_e |w_e x_e| (fp32
    rounding of a long sum is relative to the magnitude of its terms).  Returns (rows checked, max error / bound)."""
    n = g.n_rows
    rng = np.random.default_rng(seed)
    deg = (g.rowptr[1:] - g.rowptr[:-1])
    pick = np.unique(np.concatenate([[0, n - 1, int(torch.argmax(deg).item())],
                                     rng.integers(0, n, max(0, n_samples - 3))])).astype(np.int64)
    rows = torch.from_numpy(pick).to(g.device)
    lo, hi = g.rowptr[rows].cpu().numpy(), g.rowptr[rows + 1].cpu().numpy()
    lens = hi - lo
    seg = np.zeros(len(pick) + 1, dtype=np.int64)
    np.cumsum(lens, out=seg[1:])
    eidx = torch.from_numpy(np.repeat(lo - seg[:-1], lens) + np.arange(seg[-1], dtype=np.int64)).to(g.device)
    cols = g.col[eidx].long()
    w = np.ones(seg[-1]) if g.val is None else g.val[eidx].double().cpu().numpy()
    if g.col_scale is not None:
        w = w * g.col_scale[cols].double().cpu().numpy()
    uniq, inv = torch.unique(cols, return_inverse=True)
    xr = table[uniq].double().cpu().numpy()                      # only the gathered rows leave the device
    inv = inv.cpu().numpy()
    rs = np.ones(len(pick)) if g.row_scale is None else g.row_scale[rows].double().cpu().numpy()
    got = out[rows].double().cpu().numpy()
    worst = 0.0
    step = 1 << 18                                               # edges per chunk (bounds the fp64 temporaries)
    ref = np.zeros_like(got)
    mag = np.zeros_like(got)
    row_of = np.repeat(np.arange(len(pick)), lens)
    for a in range(0, int(seg[-1]), step):
        b = min(a + step, int(seg[-1]))
        t = xr[inv[a:b]] * w[a:b, None]
        np.add.at(ref, row_of[a:b], t)
        np.add.at(mag, row_of[a:b], np.abs(t))
    ref *= rs[:, None]
    mag *= np.abs(rs)[:, None]
    bound = 1e-4 * np.abs(ref) + 1e-5 * mag + 1e-30
    worst = float(np.max(np.abs(got - ref) / bound))
    return len(pick), worst


def native(args):
    import torch.distributed as dist
    from ggad_b200 import _lib, dist as gdist, ops, synth
    from ggad_b200.graph import CSRGraph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_local, m_local, d = args.nodes, args.edges, args.width
    n_glob = n_local * world
    lib = _lib.lib()

    # ---- graph shards (generated on device) ----
    t_build = time.perf_counter()
    abc = tuple(float(t) for t in args.rmat.split(",")) if args.rmat else None
    fwd = synth.rmat_shard(n_local, m_local, world, rank, seed=args.seed, device=dev, mean=True, abc=abc)
    fr = [(g * n_local, (g + 1) * n_local) for g in range(world)]
    gen = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(n_glob, d, device=dev, generator=gen)
    row_cost = None
    if world == 1:
        bwd = fwd.T
        br = fr
    else:
        # exact global in-degree of A^T rows (= out-degree histogram of sources), summed over ranks
        cnt = torch.empty(n_glob, dtype=torch.int32, device=dev)
        _lib.check(lib.ggad_col_histogram(_lib.ptr(fwd.col), fwd.nnz, _lib.ptr(cnt), n_glob, _lib.stream_ptr(dev)))
        dist.all_reduce(cnt)
        rowptr_t = torch.zeros(n_glob + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=rowptr_t[1:])
        rowptr_np = rowptr_t.cpu().numpy()
        del cnt, rowptr_t
        rs_all = torch.empty(n_glob, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(rs_all, fwd.row_scale)

        def build_bwd(rc):
            ranges = gdist.nnz_balanced_ranges(rowptr_np, world, row_cost=rc)     # balance edges + rc * rows
            lo, hi = ranges[rank]
            g = synth.rmat_transposed_shard(n_local, m_local, world, args.seed, lo, hi, device=dev, col_scale=rs_all,
                                            abc=abc).fold_col_scale()
            g.plan
            return ranges, g

        if args.row_cost == "auto":
            # profile-guided split: time the backward gather on a first split, fit t = a*nnz + b*rows over the
            # ranks (their shards differ a lot in shape: hubs vs. tail), re-split with row cost b/a
            row_cost = 2.0
            br, bwd = build_bwd(row_cost)
            for _ in range(2):
                ops.gather_reduce(bwd, x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            for _ in range(3):
                ops.gather_reduce(bwd, x)
            e1.record()
            torch.cuda.synchronize()
            samples = [None] * world
            dist.all_gather_object(samples, [bwd.n_rows, bwd.nnz, e0.elapsed_time(e1) / 3.0])
            fit = gdist.fit_row_cost(samples)
            if fit is not None and abs(fit - row_cost) > 0.15 * row_cost:
                del bwd
                row_cost = round(fit * 16) / 16.0
                br, bwd = build_bwd(row_cost)
        else:
            row_cost = float(args.row_cost)
            br, bwd = build_bwd(row_cost)
    if args.edge_order in ("hot-first", "serpentine"):
        # plan-time re-ordering of every row's edges by column popularity (CSRGraph.reorder_edges_hot_first)
        fwd = fwd.reorder_edges_hot_first(serpentine=args.edge_order == "serpentine")
        bwd = bwd.reorder_edges_hot_first(serpentine=args.edge_order == "serpentine")
    for g in (fwd, bwd):
        g.plan
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    y_full = torch.empty(n_glob, d, device=dev) if (world > 1 and args.exchange == "nccl") else None

    def compute(g, inp):
        return ops.gather_reduce(g, inp)["y"]

    # exchange of the forward output (= dY of the backward): fused into the gather epilogue over NVLink
    # peer memory (P2P stores or NVSwitch multicast), or a plain NCCL all-gather
    exchange = args.exchange if world > 1 else "none"
    rep = None
    need, halo_frac = None, None
    if exchange in ("halo", "chase", "fused", "multicast"):
        try:
            rep = gdist.PeerReplica(n_glob, d, fr, rank, dev)
            if exchange == "multicast" and not rep.multicast_ptr:
                exchange = "fused"
        except Exception as e:          # symmetric memory unavailable: fall back to the NCCL collective
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({e}); using NCCL all-gather", file=sys.stderr)
            exchange, rep = "nccl", None
    if exchange in ("halo", "chase"):
        # rows of this rank's Y shard that peer p's backward shard gathers (distinct columns of its A^T rows)
        need = gdist.halo_need_mask(bwd.col, fr, rank)
        sent = torch.zeros(1, dtype=torch.int64, device=dev)
        for s_ in range(world - 1):
            sent += ((need >> s_) & 1).sum()
        dist.all_reduce(sent)
        halo_frac = float(sent.item()) / (n_glob * (world - 1))
    peer_ptrs = rep.peer_row_ptrs if rep is not None else None
    mc_ptr = rep.multicast_row_ptr if (rep is not None and args.mc_min > 0) else None
    if args.peer_debug == "zero_mask":          # diagnostics: PEER kernel variant, nothing crosses NVLink
        need = torch.zeros(n_local, dtype=torch.int32, device=dev)
    elif args.peer_debug == "local_peers":      # diagnostics: the peer stores land in a local scratch matrix
        scratch = torch.empty(n_local, d, device=dev)
        peer_ptrs = [scratch.data_ptr()] * (world - 1)

    def step(ev=None):
        if exchange == "none":
            yy = compute(fwd, x)
            if ev:
                ev[1].record()
                ev[2].record()
        elif exchange == "nccl":
            y_loc = compute(fwd, x)
            if ev:
                ev[1].record()
            gdist.all_gather_rows(y_loc, fr, y_full)
            yy = y_full
            if ev:
                ev[2].record()
        else:
            rep.barrier(0)                       # peers finished reading the previous Y
            if exchange == "multicast":
                ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_multicast=rep.multicast_row_ptr)
            elif exchange == "chase":
                ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_peers=peer_ptrs, peer_need=need, chase=True,
                                  chase_ctas=args.chase_ctas, y_multicast=mc_ptr, mc_min_peers=args.mc_min)
            elif exchange == "halo":
                ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_peers=peer_ptrs, peer_need=need,
                                  y_multicast=mc_ptr, mc_min_peers=args.mc_min)
            else:
                ops.gather_reduce(fwd, x, y_out=rep.local_rows, y_peers=peer_ptrs, peer_need=need)
            if ev:
                ev[1].record()
            rep.barrier(1)                       # every shard has landed in every replica
            yy = rep.buf
            if ev:
                ev[2].record()
        dx_loc = compute(bwd, yy)                # dX stays sharded by source range (consumed per node)
        if ev:
            ev[3].record()
        return yy, dx_loc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    barrier()
    sampler.mark_begin()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        evs[i][0].record()
        step(evs[i])
        evs[i][4].record()
    t1.record()
    barrier()
    sampler.mark_end()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t0.elapsed_time(t1)
    tt = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    seg = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in evs])   # fwd, xchg, bwd, xchg
    seg_mean = seg.mean(0)
    shard_stats = [[bwd.n_rows, bwd.nnz]]
    if world > 1:
        st_ = [None] * world
        dist.all_gather_object(st_, [bwd.n_rows, bwd.nnz])
        shard_stats = st_
    seg_all = torch.tensor(seg_mean, dtype=torch.float64, device=dev)
    if world > 1:
        gl = [torch.empty_like(seg_all) for _ in range(world)]
        dist.all_gather(gl, seg_all)
        seg_ranks = [[round(float(v), 3) for v in t.tolist()] for t in gl]
    else:
        seg_ranks = [[round(float(v), 3) for v in seg_mean]]
    total_edges = m_local * world
    value = total_edges / (ms_per_step * 1e-3)

    # ---- parity at the measured size: sampled rows of Y and dX against an fp64 CPU recomputation, every rank ----
    verified = None
    if args.verify_rows > 0:
        yy, dx_loc = step()
        torch.cuda.synchronize()
        lo_f = fr[rank][0]
        n_y, e_y = verify_sampled_rows(fwd, x, yy[lo_f:lo_f + n_local] if yy.shape[0] == n_glob else yy, args.verify_rows, 7 + rank)
        n_dx, e_dx = verify_sampled_rows(bwd, yy, dx_loc, args.verify_rows, 11 + rank)
        halo_checked = 0
        if world > 1 and exchange != "nccl":
            # the halo rows themselves: every owner publishes a common sample of its true rows, every peer that
            # gathers one of them compares its replica bit for bit
            col_mark = torch.zeros(n_glob, dtype=torch.bool, device=dev)
            col_mark[bwd.col.long()] = True
            for g_ in range(world):
                ids = torch.from_numpy(np.random.default_rng(100 + g_).integers(fr[g_][0], fr[g_][1], 2048)).to(dev)
                truth = yy[ids].clone()
                dist.broadcast(truth, src=g_)
                if g_ != rank:
                    sel = col_mark[ids]
                    assert torch.equal(yy[ids][sel], truth[sel]), f"rank {rank}: halo rows from rank {g_} differ"
                    halo_checked += int(sel.sum().item())
            del col_mark
        stat = torch.tensor([n_y, n_dx, halo_checked, e_y, e_dx], dtype=torch.float64, device=dev)
        if world > 1:
            mx = stat.clone()
            dist.all_reduce(stat)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            stat[3:] = mx[3:]
        n_y, n_dx, halo_checked, e_y, e_dx = stat.tolist()
        verified = {"y_rows": int(n_y), "dx_rows": int(n_dx), "halo_rows_bit_exact": int(halo_checked),
                    "max_err_over_bound": max(e_y, e_dx), "bound": "1e-4*|ref| + 1e-5*sum|terms| (fp64 CPU recompute)"}
        assert max(e_y, e_dx) <= 1.0, f"parity check failed at the benchmark size: {verified}"

    clk_mhz = (clocks or {}).get("sm_mhz") if rank == 0 else None
    # ---- roofline of the dominant kernel (forward gather) ----
    peak, peak_src = hbm_peak()
    b_alg = fwd.algorithmic_bytes(d)
    fwd_s = seg_mean[0] * 1e-3
    achieved = b_alg / fwd_s / 1e9
    roofline = dict(bound="hbm", kernel="gather_tiled_kernel<G=16,CH=1,MODE=0,EPI=0,PEER=%d> (forward gather, + tile_fixup)" % int(world > 1), achieved=achieved, peak=peak,
                    unit="GB/s", frac=achieved / peak, traffic=None, peak_source=peak_src,
                    algorithmic_bytes=int(b_alg), fwd_ms=float(seg_mean[0]), bwd_ms=float(seg_mean[2]),
                    fwd_edges_per_s=m_local / fwd_s,
                    gather_bytes_model=int(fwd.nnz * (4 + 4 * d) + (fwd.n_rows + 1) * 8 + fwd.n_rows * d * 4))
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof)).get(args.workload) if world == 1 else None   # captured at N = 1 only
            if tr:
                roofline["traffic"] = tr["fwd_dram_bytes"]
                roofline["traffic_source"] = tr.get("source")
                if tr.get("fwd_l2_to_sm_bytes"):
                    # from the same ncu capture: bytes the SMs pulled out of L2 (hits + fills).  Not a ceiling: the same
                    # kernel moves 17.5 TB/s when the table is L2-resident (profiles/README.md calibration runs)
                    roofline["l2_to_sm"] = {"bytes_per_launch": tr["fwd_l2_to_sm_bytes"],
                                            "achieved_GBs": tr["fwd_l2_to_sm_bytes"] / fwd_s / 1e9,
                                            "l2_hit_rate_pct": tr.get("l2_hit_rate_pct"),
                                            "l2_resident_calibration_GBs": tr.get("l2_resident_calibration_GBs", 17500.0)}
                roofline["dram_frac_of_peak"] = tr["fwd_dram_bytes"] / fwd_s / 1e9 / peak
        except Exception:
            pass

    # ---- end-to-end: host buffers through the public entry points ----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(args, world, rank, dev, fwd, bwd, fr, br, value, rep if exchange in ("halo", "chase") else None, need)

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cn, cm = max(1000, n_local // args.cpu_frac), max(1000, m_local // args.cpu_frac)
        cpu, _, _ = cpu_reference_leg(cn, cm, d, steps=2, warmup=1)

    if rank == 0:
        out = {
            "metric": "edges/sec (SpMM fwd+bwd)", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (R-MAT %s, on-device)" % (args.rmat or "0.57/0.19/0.19/0.05"),
            "config": {"workload": args.workload, "nodes_per_gpu": n_local, "edges_per_gpu": m_local, "width": d,
                       "global_nodes": n_glob, "global_edges": total_edges, "aggregation": "mean (row_scale = 1/deg)", "edge_order": args.edge_order,
                       "l2": "inputs larger than L2 (no flush)" if n_glob * d * 4 > 2e8 else "L2-resident operand (no flush)",
                       "parallelism": f"dst-node-range x{world}, exchange={exchange}", "graph_build_s": round(t_build, 2),
                       "halo_rows_sent_frac": halo_frac, "bwd_ranges": [list(map(int, r)) for r in br], "bwd_row_cost": row_cost,
                       "bwd_shard_rows_nnz": shard_stats},
            "segments_ms": {"fwd_compute": float(seg_mean[0]), "fwd_exchange": float(seg_mean[1]),
                            "bwd_compute": float(seg_mean[2]), "bwd_exchange": float(seg_mean[3]),
                            "per_rank": seg_ranks},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "verified_rows": verified,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _bind_to_gpu_socket(index: int):
    """Pin the calling thread to the CPUs NVML reports as closest to GPU ``index`` so that the pinned host buffers
    allocated next are first-touched on that socket (8 ranks streaming 3.2 GB per step through the wrong socket's
    memory halve each other's PCIe rate).  Returns the previous affinity, or None if nothing was changed."""
    try:
        import pynvml
        old = os.sched_getaffinity(0)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            index = int(vis.split(",")[index])
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, old)
            return None
        return old
    except Exception:
        return None


def e2e_leg(args, world, rank, dev, fwd, bwd, fr, br, value, rep=None, need_y=None):
    old_affinity = _bind_to_gpu_socket(dev.index) if world > 1 else None
    try:
        return _e2e_leg(args, world, rank, dev, fwd, bwd, fr, br, value, rep, need_y)
    finally:
        if old_affinity:
            try:
                os.sched_setaffinity(0, old_affinity)
            except Exception:
                pass


def _e2e_leg(args, world, rank, dev, fwd, bwd, fr, br, value, rep=None, need_y=None):
    """Same pass with host-resident features: H2D of the rank's feature shard, D2H of its dX shard + loss."""
    import ctypes as C
    import torch.distributed as dist
    from ggad_b200 import _lib, dist as gdist, ops
    lib, ptr = _lib.lib(), _lib.ptr
    d = args.width
    n_local = args.nodes
    n_glob = n_local * world
    steps = max(2, min(args.steps, 10))
    lo_b, hi_b = br[rank]
    x_host = torch.randn(n_local, d).pin_memory()
    dx_host = torch.empty(hi_b - lo_b, d).pin_memory() if (world == 1 or rep is None) else None
    h2d, d2h = x_host.numel() * 4, (hi_b - lo_b) * d * 4 + 8
    times = []
    if world == 1:
        def res(c):
            r = _lib.ResidentCSR()
            r.rowptr, r.col, r.val = ptr(c.rowptr), ptr(c.col), ptr(c.val)
            r.row_scale, r.col_scale = ptr(c.row_scale), ptr(c.col_scale)
            r.n_rows, r.n_cols, r.nnz = c.n_rows, c.n_cols, c.nnz
            p = c.plan
            r.tile_row, r.tile_edge, r.n_tiles = ptr(p[0]), ptr(p[1]), p[2]
            return r
        ra, rt = res(fwd), res(bwd)
        # two pipeline slots (own stream + device scratch + pinned result buffers): the H2D of step i+1 overlaps
        # the D2H of step i; every step still copies its features in and its gradient + loss out
        n_slots = args.e2e_slots
        slots = []
        for _ in range(n_slots):
            slots.append(dict(stream=torch.cuda.Stream(device=dev),
                              bufs=[torch.empty(n_local, d, device=dev) for _ in range(3)],
                              ws=torch.empty(2 * max(ra.n_tiles, rt.n_tiles) * d + n_local + 16, device=dev),
                              dx=torch.empty(hi_b - lo_b, d).pin_memory(),
                              loss=torch.zeros(1, dtype=torch.float64).pin_memory()))
        torch.cuda.synchronize()

        def enqueue(i):
            s = slots[i % n_slots]
            _lib.check(lib.ggad_spmm_fwd_bwd_host_enqueue(C.byref(ra), C.byref(rt), ptr(x_host), None, ptr(s["dx"]),
                                                          ptr(s["loss"]), d, ptr(s["bufs"][0]), ptr(s["bufs"][1]),
                                                          ptr(s["bufs"][2]), ptr(s["ws"]), s["stream"].cuda_stream))
        # the host link on its own (one direction at a time), so the bound of the end-to-end figure is stated
        link = {}
        for tag, dst, src in (("h2d", slots[0]["bufs"][0], x_host), ("d2h", slots[0]["dx"], slots[0]["bufs"][2])):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            link[tag + "_GBs_alone"] = round(3 * src.numel() * 4 / (time.perf_counter() - t0) / 1e9, 1)
        # ... and both directions at once on two streams: the ceiling of a step that overlaps its copies perfectly
        s_in, s_out = slots[0]["stream"], torch.cuda.Stream(device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s_in):
                slots[0]["bufs"][0].copy_(x_host, non_blocking=True)
            with torch.cuda.stream(s_out):
                slots[0]["dx"].copy_(slots[0]["bufs"][2], non_blocking=True)
        torch.cuda.synchronize()
        both = 3 * (x_host.numel() + slots[0]["dx"].numel()) * 4 / (time.perf_counter() - t0) / 1e9
        link["both_GBs_concurrent"] = round(both, 1)
        for i in range(n_slots):                # warm-up
            enqueue(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            enqueue(i)
        torch.cuda.synchronize()
        times = [(time.perf_counter() - t0) / steps]
        assert all(float(s["loss"][0]) > 0 for s in slots)
        api = f"ggad_spmm_fwd_bwd_host_enqueue (C ABI, pinned host buffers, {n_slots} pipelined streams)"
    elif rep is not None:
        # N > 1, halo exchange end to end: every step copies the rank's feature shard in from pinned host memory,
        # pushes the rows its peers' forward shards gather (ggad_halo_push), runs the fused forward (+ halo push of
        # Y) and the backward, and copies its dX shard + loss out.  The H2D of step i+1 runs on a second stream
        # into a staging buffer, so it overlaps the D2H of step i (PCIe is full duplex).
        x_rep = gdist.PeerReplica(n_glob, d, fr, rank, dev)
        need_x = gdist.halo_need_mask(fwd.col, fr, rank)
        cur = torch.cuda.current_stream(dev)
        h_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty(n_local, d, device=dev) for _ in range(2)]
        # dX leaves through the node's owner: the backward ranges are compute-balanced (rank 0 owns few hub rows,
        # the last rank most of the graph), so the shards are re-sharded to the even node ranges over NVLink first
        # and every rank copies out exactly n_local rows
        dx_own = gdist.PeerBlock(n_local, d, dev)
        dx_pin = [torch.empty(n_local, d).pin_memory() for _ in range(2)]
        d2h = n_local * d * 4 + 4
        loss_pin = [torch.zeros(1).pin_memory() for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(h_stream):
                if i >= 2:
                    h_stream.wait_event(ev_free[i % 2])
                stage[i % 2].copy_(x_host, non_blocking=True)
                ev_in[i % 2].record(h_stream)

        def run(i):
            cur.wait_event(ev_in[i % 2])
            x_rep.local_rows.copy_(stage[i % 2])
            ev_free[i % 2].record(cur)
            x_rep.barrier(0)                    # every rank is done gathering the previous X
            ops.halo_push(x_rep.local_rows, x_rep.peer_row_ptrs, need_x)
            x_rep.barrier(1)                    # all X halos have landed (and every rank finished its last backward)
            r = ops.gather_reduce(fwd, x_rep.buf, y_out=rep.local_rows, y_peers=rep.peer_row_ptrs, peer_need=need_y,
                                  want_sumsq=True)
            rep.barrier(1)                      # all Y halos have landed
            dx = ops.gather_reduce(bwd, rep.buf)["y"]
            dx_own.barrier(0)                   # the owners' previous dX blocks have been copied out
            gdist.reshard_rows(dx, (lo_b, hi_b), fr, dx_own)
            dx_own.barrier(1)
            dx_pin[i % 2].copy_(dx_own.buf, non_blocking=True)
            loss_pin[i % 2].copy_(r["sumsq"].sum().mul_(0.5).reshape(1), non_blocking=True)

        warm = 2
        prefetch(0)
        for i in range(warm):
            prefetch(i + 1)
            run(i)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(warm, warm + steps):
            if i + 1 < warm + steps:
                prefetch(i + 1)
            run(i)
        torch.cuda.synchronize()
        times = [(time.perf_counter() - t0) / steps]
        assert all(float(l[0]) > 0 for l in loss_pin)
        api = ("ggad_b200.ops.halo_push + gather_reduce (fused halo exchange), pinned host shards in/out, "
               "H2D of step i+1 overlapped with D2H of step i")
    else:
        x_full = torch.empty(n_glob, d, device=dev)
        y_full = torch.empty(n_glob, d, device=dev)
        x_loc = torch.empty(n_local, d, device=dev)
        for i in range(steps + 1):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x_loc.copy_(x_host, non_blocking=True)
            gdist.all_gather_rows(x_loc, fr, x_full)
            r = ops.gather_reduce(fwd, x_full, want_sumsq=True)
            gdist.all_gather_rows(r["y"], fr, y_full)
            dx = ops.gather_reduce(bwd, y_full)["y"]
            dx_host.copy_(dx, non_blocking=True)
            loss = 0.5 * float(r["sumsq"].double().sum().item())
            torch.cuda.synchronize()
            if i:
                times.append(time.perf_counter() - t0)
        api = "ggad_b200.ops.gather_reduce + dist.all_gather_rows (pinned host shards)"
    t = torch.tensor([float(np.mean(times))], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    res = {"value": args.edges * world / sec, "unit": "edges/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": sec * 1e3, "steps": steps, "api": api,
           "host_link_GBs_achieved": round((h2d + d2h) / sec / 1e9, 1),                 # per rank, both directions
           "host_link_GBs_aggregate": round(world * (h2d + d2h) / sec / 1e9, 1)}        # the box's host memory serves all ranks
    if world == 1:
        res["host_link"] = link
    return res


def reference(args):
    """Reference arm: the reference's own CPU path for the op on the SAME config as the native arm (full per-GPU
    workload: S64 = 6.25 M nodes / 125 M edges / d = 64, ~7 s per step on 16 threads).  K is honoured up to a wall
    budget (--ref-budget seconds of timed steps, default 150; at least 3 steps), the JSON reports the steps run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_local, m_local, d = args.nodes, args.edges, args.width
    cn, cm = max(1000, n_local // args.cpu_frac), max(1000, m_local // args.cpu_frac)
    warm = max(1, min(args.warmup, 1 if cm > 20_000_000 else 2))
    cpu, t, done = cpu_reference_leg(cn, cm, d, steps=max(1, args.steps), warmup=warm, budget_s=args.ref_budget)
    out = {
        "impl": "reference", "metric": "edges/sec (SpMM fwd+bwd)", "value": cpu["value"], "unit": "edges/s",
        "n_gpus": args.gpus, "steps": done, "warmup": warm, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (R-MAT 0.57/0.19/0.19/0.05, host numpy)",
        "config": {"workload": args.workload, "nodes_per_gpu": n_local, "edges_per_gpu": m_local, "width": d,
                   "same_config": args.cpu_frac == 1, "steps_requested": args.steps, "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def minibatch_workload(args):
    """--workload C4mb: one step = one training batch of program B (150 + 50 nodes: device frontier, two gathers,
    fused dense tail, backward, Adam replayed as a CUDA graph -- train.GraphedMiniBatchStep; src/model_handler.py:330-364)
    on the DGraph-shaped synthetic graph; N ranks = data parallel replicas with their own seed batches.  Re-emits tools/bench_minibatch.py's result in the bench contract."""
    import io
    import contextlib
    import runpy
    rank = int(os.environ.get("RANK", "0"))
    argv = ["bench_minibatch.py", "--nodes", str(args.nodes), "--edges", str(args.edges), "--d", str(args.width),
            "--iters", str(max(10, args.steps)), "--warm", str(max(3, args.warmup)), "--graphed", "--no-prefetch"]
    argv += ["--cpu-nodes", "0"] if (args.no_cpu or args.impl != "reference") else []
    old = sys.argv
    sys.argv = argv
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(ROOT, "tools", "bench_minibatch.py"), run_name="__main__")
    finally:
        sys.argv = old
    if rank != 0:
        return
    r = json.loads(buf.getvalue().strip().splitlines()[-1])
    emit({"metric": "mini-batches/sec (GGAD GraphSAGE training batch, 150 + 50 nodes)", "value": r["batches_per_s"],
          "unit": "batches/s", "n_gpus": r["n_gpus"], "steps": max(10, args.steps), "warmup": max(3, args.warmup),
          "ms_per_step": r["wall_ms_per_lockstep_batch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic (DGraph-shaped R-MAT, on-device)",
          "config": {"workload": "C4mb", "nodes": r["nodes"], "adjacency_entries": r["adjacency_entries"], "width": r["d"],
                     "h": r["h"], "batch": r["batch"], "parallelism": f"data parallel x{r['n_gpus']} (replicated graph + table)"},
          "detail": r, "gpu_launches": int(r["ggad_launches_per_batch"] * max(10, args.steps))})


_RESULT_FD = None


def _reserve_stdout():
    """stdout carries exactly ONE line, the result JSON: anything a library prints there (NCCL's version banner, torch
    warnings) is sent to stderr by pointing fd 1 at fd 2; emit() writes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, line)
    else:
        os.write(_RESULT_FD, line)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="S64", choices=list(WORKLOADS) + ["custom"])
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--edges", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--rmat", default=None, help="a,b,c of the R-MAT generator (default 0.57,0.19,0.19); 0.25,0.25,0.25 = uniform")
    ap.add_argument("--cpu-frac", type=int, default=None,
                    help="CPU legs run on 1/frac of the per-GPU workload (default: 8 for the native arm's cpu_baseline "
                         "sample, 1 = the full config for --impl reference)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: wall budget of the timed steps (s)")
    ap.add_argument("--chase-ctas", type=int, default=0, help="exchange=chase: CTAs of the chase kernel (0 = library default)")
    ap.add_argument("--mc-min", type=int, default=0,
                    help="exchange=halo|chase: rows needed by at least this many peers go once through the NVSwitch multicast "
                         "address instead of one unicast store per peer (0 = never)")
    ap.add_argument("--verify-rows", type=int, default=4096, help="rows of Y and of dX recomputed on the CPU after the timed region")
    ap.add_argument("--exchange", default="halo", choices=["chase", "halo", "fused", "multicast", "nccl"],
                    help="N>1: how the forward output reaches the ranks that gather it next.  chase = the gather kernel "
                         "flags finished tiles and a concurrent kernel on a few SMs stores the halo rows over NVLink; "
                         "halo = NVLink P2P stores "
                         "from the gather epilogue, only rows a peer's next pass reads; fused = same, every row to every "
                         "peer; multicast = one NVSwitch multimem.st per row; nccl = separate all-gather")
    ap.add_argument("--row-cost", default="auto",
                    help="N>1: cost of one output row in edges when the backward ranges are balanced; 'auto' fits it from "
                         "a timed trial split (profile-guided)")
    ap.add_argument("--peer-debug", default=None, choices=["zero_mask", "local_peers"],
                    help="diagnostics only (results are NOT exchanged): isolate the cost of the peer stores")
    ap.add_argument("--edge-order", default="column", choices=["column", "hot-first", "serpentine"],
                    help="order of the edges inside a CSR row: by column id (as a CSR build leaves them) or most popular "
                         "column first (homogeneous L2-hit / DRAM-miss batches)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slots", type=int, default=3, help="N=1 end-to-end leg: pipeline depth (streams with their own scratch)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.workload != "custom":
        n, m, d = WORKLOADS[args.workload]
        args.nodes = args.nodes or n
        args.edges = args.edges or m
        args.width = args.width or d
    if args.workload == "C4mb":
        return minibatch_workload(args)
    if args.impl == "reference":
        args.cpu_frac = args.cpu_frac or 1
        reference(args)
    else:
        args.cpu_frac = args.cpu_frac or 8
        native(args)


if __name__ == "__main__":
    main()
