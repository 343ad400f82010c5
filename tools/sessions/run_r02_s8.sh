#!/bin/bash
# Round 2, GPU session 8 (2 GPUs): multi-GPU tests with oracle diagnostics; program-A epoch launch list (regression hunt);
# program B graphed with the larger capacities, with / without the block prefetch thread.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02i_pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -n "AssertionError\|passed\|failed" $O/r02i_pytest_multi.log | cut -c1-1800
export CUDA_VISIBLE_DEVICES=0
timeout 300 python tools/bench_dense.py > $O/r02i_dense.txt 2>&1; cat $O/r02i_dense.txt
for c in C1 C2 C3; do timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> $O/r02i_epoch.jsonl 2>> $O/r02i_epoch.err; done
python - <<PY
import json
for ln in open("$O/r02i_epoch.jsonl"):
    j=json.loads(ln); print(j["config"], "eager", round(j["gpu_epoch_ms_events"],3), "graph", round(j["cuda_graph_epoch_ms_events"],3))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02i_launches_epoch_C2.csv python tools/bench_epoch.py --config C2 --epochs 2 --cpu-epochs 0 --no-graph > $O/r02i_launches_epoch_C2.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/r02i_launches_epoch_C2.csv 2>/dev/null | head -45
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed > $O/r02i_minibatch_graphed_pf.json 2> $O/r02i_minibatch_graphed_pf.err; echo "mb graphed prefetch exit $?"; cut -c1-420 $O/r02i_minibatch_graphed_pf.json
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02i_minibatch_graphed.json 2> $O/r02i_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-420 $O/r02i_minibatch_graphed.json
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 > $O/r02i_minibatch_pf.json 2> $O/r02i_minibatch_pf.err; echo "mb eager prefetch exit $?"; cut -c1-420 $O/r02i_minibatch_pf.json
