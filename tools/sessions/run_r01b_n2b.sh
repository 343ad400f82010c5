#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/n2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2b_pytest.log
tail -8 gpurun_out/n2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
run() { # name, extra args...
  name=$1; shift
  timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/n2b_$name.json 2> gpurun_out/n2b_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/n2b_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"],3), round(j["value"]/1e9,2), j["segments_ms"]["per_rank"], j["config"].get("halo_rows_sent_frac"))
except Exception as e:
    print("$name", "parse failed", e)
PY
}
run halo_zero --exchange halo --peer-debug zero_mask
run halo --exchange halo
run fused --exchange fused
run multicast --exchange multicast
run nccl --exchange nccl
