#!/bin/bash
# Round 2, last 1-GPU sanity run on the shipped library: suite, smoke, default bench line, program B graphed.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r02zz}
timeout 1200 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/${T}_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${T}_smoke.log
timeout 600 python bench.py > $O/${T}_bench_S64_n1.json 2> $O/${T}_bench_S64_n1.err; echo "bench exit $?"; python - <<PY
import json
j=json.loads(open("$O/${T}_bench_S64_n1.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e9,2), "GE/s", round(j["ms_per_step"],3), "ms  roofline", round(j["roofline"]["frac"],3), "traffic", j["roofline"]["traffic"], "e2e", round(j["e2e"]["ms_per_step"],2), "cpu", round(j["cpu_baseline"]["value"]/1e6,1), "launches", j["gpu_launches"], j["verified_rows"], j["clocks"])
PY
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/${T}_minibatch_graphed.json 2> $O/${T}_minibatch_graphed.err; cut -c1-300 $O/${T}_minibatch_graphed.json
