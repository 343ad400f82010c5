#!/bin/bash
# Round 2, GPU session 6 (1 GPU): suite after split-K / stream-K, static-shape mini-batch loss, block prefetch, lower
# plan threshold; program-A epochs (regression check vs r01), program-B batch with and without the input pipeline.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02f_pytest.log 2>&1; echo "pytest exit $?"; tail -25 $O/r02f_pytest.log | cut -c1-900
timeout 300 python tools/bench_dense.py > $O/r02f_dense.txt 2>&1; cat $O/r02f_dense.txt
for c in C1 C2 C3; do timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> $O/r02f_epoch.jsonl 2>> $O/r02f_epoch.err; done; echo "epoch exit $?"
python - <<PY
import json
for ln in open("$O/r02f_epoch.jsonl"):
    j=json.loads(ln); print(j["config"], "eager", round(j["gpu_epoch_ms_events"],3), "graph", round(j["cuda_graph_epoch_ms_events"],3), "launches", j["ggad_launches_per_epoch"])
PY
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 > $O/r02f_minibatch_prefetch.json 2> $O/r02f_minibatch_prefetch.err; echo "mb prefetch exit $?"; cut -c1-900 $O/r02f_minibatch_prefetch.json
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/r02f_minibatch_noprefetch.json 2> $O/r02f_minibatch_noprefetch.err; echo "mb no-prefetch exit $?"; cut -c1-600 $O/r02f_minibatch_noprefetch.json
timeout 300 python tools/bench_minibatch.py --phases --cpu-nodes 0 --iters 60 > $O/r02f_minibatch_phases.json 2> $O/r02f_minibatch_phases.err; python -c "
import json; print(json.loads(open('$O/r02f_minibatch_phases.json').read())['phase_ms_mean'])"
