#!/bin/bash
# Round 2, GPU session 15 (1 GPU): fused mini-batch tail kernels -- parity (goldens, torch formulation, graphed step), program B
# batch time eager / graphed, launch list of the graphed batch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_kernels.py -m gpu -q > $O/r02r_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02r_pytest.log | cut -c1-300; grep -n "Error\|FAILED\|^E " $O/r02r_pytest.log | head -12 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02r_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r02r_smoke.log
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02r_minibatch_graphed.json 2> $O/r02r_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-330 $O/r02r_minibatch_graphed.json; tail -2 $O/r02r_minibatch_graphed.err
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/r02r_minibatch_eager.json 2> $O/r02r_minibatch_eager.err; cut -c1-330 $O/r02r_minibatch_eager.json
GGAD_TORCH_TAIL=1 timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02r_minibatch_graphed_torchtail.json 2> /dev/null; cut -c1-330 $O/r02r_minibatch_graphed_torchtail.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 600 --csv --log-file $O/r02r_launches_minibatch_graphed.csv python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 6 --warm 12 > $O/r02r_launches.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/r02r_launches_minibatch_graphed.csv 2>/dev/null | head -24
