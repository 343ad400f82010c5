#!/bin/bash
# Round 2, GPU session 10 (1 GPU): suite after the hop-2 direct block; program B eager / graphed, phase timers, launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02l_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02l_pytest.log | cut -c1-600; grep -n "Error\|assert" $O/r02l_pytest.log | head -10 | cut -c1-400
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02l_minibatch_graphed.json 2> $O/r02l_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-420 $O/r02l_minibatch_graphed.json; tail -2 $O/r02l_minibatch_graphed.err
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed > $O/r02l_minibatch_graphed_pf.json 2> $O/r02l_minibatch_graphed_pf.err; echo "mb graphed prefetch exit $?"; cut -c1-420 $O/r02l_minibatch_graphed_pf.json
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/r02l_minibatch_eager.json 2> $O/r02l_minibatch_eager.err; echo "mb eager exit $?"; cut -c1-420 $O/r02l_minibatch_eager.json
timeout 300 python tools/bench_minibatch.py --phases --cpu-nodes 0 --iters 60 > $O/r02l_minibatch_phases.json 2> $O/r02l_minibatch_phases.err; python -c "
import json; print(json.loads(open('$O/r02l_minibatch_phases.json').read())['phase_ms_mean'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1000 --csv --log-file $O/r02l_launches_minibatch_graphed.csv python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 6 --warm 12 > $O/r02l_launches_minibatch_graphed.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/r02l_launches_minibatch_graphed.csv 2>/dev/null | head -30
