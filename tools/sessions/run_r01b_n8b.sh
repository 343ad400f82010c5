#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # n, name, extra args...
  n=$1; name=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/n8b_$name.json 2> gpurun_out/n8b_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/n8b_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"],3), round(j["value"]/1e9,2), j["segments_ms"]["per_rank"], j["config"].get("halo_rows_sent_frac"), j["config"]["bwd_row_cost"], j["config"]["bwd_shard_rows_nnz"], j["e2e"], j["clocks"])
except Exception as e:
    print("$name", "parse failed", e)
PY
  tail -2 gpurun_out/n8b_$name.err
}
run 8 halo8
run 4 halo4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 tools/bench_minibatch.py --iters 60 > gpurun_out/n8b_minibatch.json 2> gpurun_out/n8b_minibatch.err; echo "mb8 exit $?"; cat gpurun_out/n8b_minibatch.json; tail -2 gpurun_out/n8b_minibatch.err
