#!/bin/bash
# Round 2, GPU session 11 (1 GPU): suite after the padded-plan graphed step and the column-subset affinity backward;
# program B graphed; program A epochs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02m_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02m_pytest.log | cut -c1-600; grep -n "Error\|FAILED" $O/r02m_pytest.log | head -10 | cut -c1-400
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02m_minibatch_graphed.json 2> $O/r02m_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-420 $O/r02m_minibatch_graphed.json; tail -2 $O/r02m_minibatch_graphed.err
for c in C1 C2 C3; do timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> $O/r02m_epoch.jsonl 2>> $O/r02m_epoch.err; done
python - <<PY
import json
for ln in open("$O/r02m_epoch.jsonl"):
    j=json.loads(ln); print(j["config"], "eager", round(j["gpu_epoch_ms_events"],3), "graph", round(j["cuda_graph_epoch_ms_events"],3))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1000 --csv --log-file $O/r02m_launches_minibatch_graphed.csv python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 6 --warm 12 > $O/r02m_launches_minibatch_graphed.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/r02m_launches_minibatch_graphed.csv 2>/dev/null | head -14
