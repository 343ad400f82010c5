#!/bin/bash
# Round 2, GPU session 13 (1 GPU): edges inside a row ordered most-popular-column-first (homogeneous L2-hit / DRAM-miss
# batches) against the column-id order, on S64, C4 and uniform columns; parity test of the re-ordering.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "hot_first or edge_cases" > $O/r02p_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/r02p_pytest.log | cut -c1-300
for wl in S64 C4; do for eo in column hot-first; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-cpu --edge-order $eo > $O/r02p_bench_${wl}_$eo.json 2> $O/r02p_bench_${wl}_$eo.err
  python - <<PY
import json
j=json.loads(open("$O/r02p_bench_${wl}_$eo.json").read().strip().splitlines()[-1])
print("$wl $eo", "fwd", round(j["segments_ms"]["fwd_compute"],3), "bwd", round(j["segments_ms"]["bwd_compute"],3), "GE/s", round(j["value"]/1e9,2), "verify", round(j["verified_rows"]["max_err_over_bound"],4), "build_s", j["config"]["graph_build_s"])
PY
done; done
timeout 600 ncu --set full --clock-control none -k regex:gather_tiled_kernel -s 2 -c 2 -f -o $O/r02p_ncu_S64_hotfirst python tools/profile_spmm.py --workload S64 --iters 2 --variant hot_first > $O/r02p_ncu.log 2>&1; echo "ncu exit $?"
