#!/bin/bash
# Round 2, GPU session 4 (2 GPUs): whole suite again (multi-GPU oracle message, sharded-feature tests, f3/f4 tests),
# C5-shaped sharded-feature SAGE bench at N=2, N=2 S64 bench with the default exchange.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02d_pytest.log 2>&1; echo "pytest exit $?"; tail -25 $O/r02d_pytest.log | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR tools/bench_sharded_sage.py --mode both > $O/r02d_sharded_sage_n2.json 2> $O/r02d_sharded_sage_n2.err; echo "sharded sage exit $?"; tail -c 1500 $O/r02d_sharded_sage_n2.json; tail -5 $O/r02d_sharded_sage_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/bench_sharded_sage.py --mode both > $O/r02d_sharded_sage_n1.json 2> $O/r02d_sharded_sage_n1.err; echo "sharded sage n1 exit $?"; tail -c 1000 $O/r02d_sharded_sage_n1.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02d_n2_bench.json 2> $O/r02d_n2_bench.err; echo "bench exit $?"; tail -c 600 $O/r02d_n2_bench.json
