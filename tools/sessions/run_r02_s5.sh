#!/bin/bash
# Round 2, GPU session 5 (1 GPU): where a program-B batch spends its time (phase timers + ncu launch list), the
# N=1 bench with the 3-slot end-to-end pipeline and host-link figures, the C4mb wrapper.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/bench_minibatch.py --phases --cpu-nodes 0 --iters 60 > $O/r02e_minibatch_phases.json 2> $O/r02e_minibatch_phases.err; echo "phases exit $?"; tail -c 1200 $O/r02e_minibatch_phases.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file $O/r02e_launches_minibatch.csv python tools/bench_minibatch.py --cpu-nodes 0 --iters 6 --warm 12 > $O/r02e_launches_minibatch.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/r02e_launches_minibatch.csv 2>/dev/null | head -40
for sl in 2 3 4; do timeout 600 python bench.py --steps 20 --warmup 5 --e2e-slots $sl --no-cpu > $O/r02e_bench_slots$sl.json 2> $O/r02e_bench_slots$sl.err; echo "bench slots $sl exit $?"; python - <<PY
import json
j=json.loads(open("$O/r02e_bench_slots$sl.json").read().strip().splitlines()[-1])
print($sl, "e2e", j["e2e"])
PY
done
timeout 600 python bench.py --workload C4mb --steps 50 --warmup 10 > $O/r02e_bench_C4mb.json 2> $O/r02e_bench_C4mb.err; echo "C4mb exit $?"; cut -c1-700 $O/r02e_bench_C4mb.json
