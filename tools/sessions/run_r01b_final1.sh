#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f1_pytest.log
tail -3 gpurun_out/f1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/f1_smoke.log
timeout 600 python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/f1_bench.json; tail -3 gpurun_out/f1_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err; echo "ref exit $?"; tail -c 800 gpurun_out/f1_bench_ref.json
timeout 300 python tools/bench_minibatch.py --iters 60 --cpu-nodes 0 > gpurun_out/f1_minibatch.json 2> gpurun_out/f1_minibatch.err; echo "mb exit $?"; cat gpurun_out/f1_minibatch.json
