#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # n, name, extra args...
  n=$1; name=$2; shift; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/n8_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"],3), round(j["value"]/1e9,2), j["segments_ms"]["per_rank"], j["config"].get("halo_rows_sent_frac"), j["config"]["bwd_shard_rows_nnz"])
except Exception as e:
    print("$name", "parse failed", e)
PY
}
run 8 halo8 --exchange halo
run 8 fused8 --exchange fused
run 8 nccl8 --exchange nccl
run 4 halo4 --exchange halo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/p2p_probe.py > gpurun_out/n8_p2p_probe.json 2> gpurun_out/n8_p2p_probe.err; cat gpurun_out/n8_p2p_probe.json
