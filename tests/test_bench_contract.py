"""bench.py contract checks that run without a GPU: the reference arm prints ONE JSON line with the agreed keys,
non-zero ranks of a multi-process launch stay silent, and the host R-MAT generator used by the CPU legs has the
same shape as the device generator (row-sorted CSR, columns in range, duplicate edges kept)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "custom",
                        "--nodes", "20000", "--edges", "200000", "--width", "16", "--cpu-frac", "1", "--steps", "1",
                        "--warmup", "1", *extra], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_contract():
    out = _run([])
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "edges/sec (SpMM fwd+bwd)" and j["unit"] == "edges/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["dtype"] == "f32"
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["config"]["workload"] == "custom" and j["config"]["width"] == 16
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "torch.spmm" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    assert _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_host_rmat_generator_shape():
    sys.path.insert(0, ROOT)
    import bench
    n, m = 5000, 60000
    rowptr, col = bench.rmat_csr_numpy(n, m, seed=0)
    assert rowptr.dtype == np.int64 and col.dtype == np.int32 and len(rowptr) == n + 1
    assert rowptr[0] == 0 and rowptr[-1] == m and np.all(np.diff(rowptr) >= 0)
    assert col.min() >= 0 and col.max() < n
    for r in (0, 1, 17, n - 1):                                      # columns sorted inside a row
        seg = col[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(seg) >= 0)
    deg = np.diff(rowptr)
    assert deg.max() > 20 * deg.mean()                               # power-law head (R-MAT 0.57/0.19/0.19/0.05)
    rp2, col2 = bench.rmat_csr_numpy(n, m, seed=0)
    assert np.array_equal(rowptr, rp2) and np.array_equal(col, col2)  # deterministic


def test_clock_sampler_degrades_without_nvml():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_in_bench_row_verification_catches_a_wrong_row():
    """bench.verify_sampled_rows (the parity check bench.py runs after the timed region at the benchmark's own size):
    passes on a correct fp32 SpMM result, fails on a single perturbed row; works on any object with CSR attributes."""
    import types
    import scipy.sparse as sp
    import torch
    sys.path.insert(0, ROOT)
    import bench
    n, d = 20000, 16
    rp, col = bench.rmat_csr_numpy(n, 300000, 3)
    deg = np.diff(rp)
    rng = np.random.default_rng(0)
    g = types.SimpleNamespace(rowptr=torch.from_numpy(rp), col=torch.from_numpy(col), val=torch.from_numpy(rng.random(len(col)).astype(np.float32)),
                              n_rows=n, row_scale=torch.from_numpy(np.where(deg > 0, 1.0 / np.maximum(deg, 1), 0).astype(np.float32)),
                              col_scale=torch.from_numpy(rng.random(n).astype(np.float32)), device=torch.device("cpu"))
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    a = sp.csr_matrix((g.val.numpy() * g.col_scale.numpy()[col], col, rp), shape=(n, n))
    y = torch.from_numpy((sp.diags(g.row_scale.numpy()) @ a @ x.numpy()).astype(np.float32))
    rows, worst = bench.verify_sampled_rows(g, x, y, 1024, 0)
    assert rows > 900 and worst < 0.5
    hub = int(np.argmax(deg))                       # the highest-degree row is always in the sample
    y[hub] += 1e-2 * (1.0 + y[hub].abs())
    assert bench.verify_sampled_rows(g, x, y, 1024, 0)[1] > 1.0
