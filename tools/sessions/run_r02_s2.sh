#!/bin/bash
# Round 2, GPU session 2 (1 GPU): K6 dense projections (tcgen05 GEMM) parity, restructured in-kernel push and chase
# kernel v2 on S64 with local stand-in peers, per-group push for comparison.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest exit $?"; tail -15 $O/r02b_pytest.log
timeout 300 python tools/bench_variants.py --workload S64 > $O/r02b_variants_S64_shipped.txt 2>&1; echo "variants exit $?"; cat $O/r02b_variants_S64_shipped.txt
GGAD_B200_LIB=$PWD/ab_push_group.so timeout 300 python tools/bench_variants.py --workload S64 --no-chase > $O/r02b_variants_S64_ab_push_group.txt 2>&1; echo "variants push_group exit $?"
grep -i "push\|plain" $O/r02b_variants_S64_ab_push_group.txt
timeout 300 python tools/bench_dense.py > $O/r02b_dense.txt 2>&1; echo "dense exit $?"; cat $O/r02b_dense.txt
for c in C1 C2 C3; do timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> $O/r02b_epoch.jsonl 2>> $O/r02b_epoch.err; done; echo "epoch exit $?"; tail -c 1500 $O/r02b_epoch.jsonl
timeout 300 python tools/bench_minibatch.py > $O/r02b_minibatch.json 2> $O/r02b_minibatch.err; echo "minibatch exit $?"; tail -c 800 $O/r02b_minibatch.json
