#!/bin/bash
# Round 2, GPU session 12 (1 GPU): suite on the cleaned-up library (deferred epilogue kind removed, hybrid multicast in the
# push, CTA-per-row block fill), evaluation bench (f3), program B graphed.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02o_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02o_pytest.log | cut -c1-300; grep -n "Error\|FAILED" $O/r02o_pytest.log | head -8 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02o_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r02o_smoke.log
timeout 600 python tools/bench_eval.py > $O/r02o_eval.json 2> $O/r02o_eval.err; echo "eval exit $?"; cat $O/r02o_eval.json; tail -2 $O/r02o_eval.err
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02o_minibatch_graphed.json 2> $O/r02o_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-330 $O/r02o_minibatch_graphed.json
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/r02o_minibatch_eager.json 2> $O/r02o_minibatch_eager.err; cut -c1-330 $O/r02o_minibatch_eager.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02o_bench_S64_n1.json 2> $O/r02o_bench_S64_n1.err; echo "bench exit $?"; cut -c1-330 $O/r02o_bench_S64_n1.json
