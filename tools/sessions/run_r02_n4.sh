#!/bin/bash
# Round 2, 4-GPU session: S64 x 4 layer pass (scaling table), multi-GPU tests at world 4, C5-shaped sharded SAGE and
# program B data parallel (graphed tail) at N=4.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02k_pytest_multi_n4.log 2>&1; echo "pytest multi exit $?"; grep -n "AssertionError\|passed\|failed" $O/r02k_pytest_multi_n4.log | cut -c1-1500
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02k_n${N}_halo.json 2> $O/r02k_n${N}_halo.err; echo "bench halo exit $?"
python - <<PY
import json
j=json.loads(open("$O/r02k_n${N}_halo.json").read().strip().splitlines()[-1])
print("halo", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", [r[:3:2] for r in j["segments_ms"]["per_rank"]], "e2e", round(j["e2e"]["ms_per_step"],1), j["config"]["halo_rows_sent_frac"], j["verified_rows"])
PY
timeout 900 $TR tools/bench_sharded_sage.py --mode both --iters 20 --warm 5 > $O/r02k_sharded_sage_n${N}.json 2> $O/r02k_sharded_sage_n${N}.err; echo "sharded sage exit $?"; tail -c 1500 $O/r02k_sharded_sage_n${N}.json
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02k_minibatch_graphed_n${N}.json 2> $O/r02k_minibatch_graphed_n${N}.err; echo "minibatch graphed exit $?"; cut -c1-500 $O/r02k_minibatch_graphed_n${N}.json
