#!/bin/bash
# Round 2, GPU session 14 (2 GPUs): multi-GPU tests incl. the hybrid multicast rule inside the in-kernel push, N=2 bench with
# the default exchange and with --mc-min 1, serpentine edge order on one GPU.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -q -k "nccl or hot_first" > $O/r02q_pytest.log 2>&1; echo "pytest exit $?"; grep -n "AssertionError\|passed\|failed" $O/r02q_pytest.log | cut -c1-700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
for mc in 0 1; do
  timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --mc-min $mc --no-cpu --no-e2e > $O/r02q_n2_halo_mc$mc.json 2> $O/r02q_n2_halo_mc$mc.err; echo "bench mc$mc exit $?"
  python - <<PY
import json
try:
    j=json.loads(open("$O/r02q_n2_halo_mc$mc.json").read().strip().splitlines()[-1])
    print("mc_min=$mc", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", j["segments_ms"]["per_rank"], j["verified_rows"]["halo_rows_bit_exact"])
except Exception as e: print("mc_min=$mc failed", e)
PY
done
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --edge-order serpentine > $O/r02q_bench_S64_serpentine.json 2> $O/r02q_bench_S64_serpentine.err
python - <<PY
import json
j=json.loads(open("$O/r02q_bench_S64_serpentine.json").read().strip().splitlines()[-1])
print("S64 serpentine", "fwd", round(j["segments_ms"]["fwd_compute"],3), "bwd", round(j["segments_ms"]["bwd_compute"],3), "verify", round(j["verified_rows"]["max_err_over_bound"],4))
PY
