#!/bin/bash
# Round 2, final 2-GPU check: whole suite incl. the multi-GPU tests, default bench at N=2 (with the end-to-end leg),
# program B data parallel with the graphed fused tail.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02y_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02y_pytest.log | cut -c1-300; grep -n "FAILED\|^E " $O/r02y_pytest.log | head -10 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02y_bench_S64_n2.json 2> $O/r02y_bench_S64_n2.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$O/r02y_bench_S64_n2.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", j["segments_ms"]["per_rank"], "e2e", j["e2e"]["ms_per_step"], j["e2e"].get("host_link_GBs_aggregate"), j["verified_rows"], j["clocks"])
PY
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02y_minibatch_graphed_n2.json 2> $O/r02y_minibatch_graphed_n2.err; echo "mb graphed n2 exit $?"; cut -c1-330 $O/r02y_minibatch_graphed_n2.json
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/r02y_minibatch_eager_n2.json 2> $O/r02y_minibatch_eager_n2.err; echo "mb eager n2 exit $?"; cut -c1-330 $O/r02y_minibatch_eager_n2.json
