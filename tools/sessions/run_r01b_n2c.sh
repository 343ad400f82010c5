#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -x -q -k "exchange or sharded" > gpurun_out/n2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2c_pytest.log
tail -4 gpurun_out/n2c_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_variants.py --workload S64 > gpurun_out/n2c_variants.txt 2>&1; cat gpurun_out/n2c_variants.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
run() { # name, extra args...
  name=$1; shift
  timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/n2c_$name.json 2> gpurun_out/n2c_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/n2c_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"],3), round(j["value"]/1e9,2), j["segments_ms"]["per_rank"], j["config"].get("halo_rows_sent_frac"), j["config"]["bwd_shard_rows_nnz"])
except Exception as e:
    print("$name", "parse failed", e)
PY
}
run halo_zero --exchange halo --peer-debug zero_mask --row-cost 1
run halo_rc1 --exchange halo --row-cost 1
run halo_rc2 --exchange halo --row-cost 2
run halo_rc3 --exchange halo --row-cost 3
run halo_rc4 --exchange halo --row-cost 4
run fused_rc2 --exchange fused --row-cost 2
