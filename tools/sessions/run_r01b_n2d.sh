#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2e_pytest.log
tail -6 gpurun_out/n2e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2e_bench.json 2> gpurun_out/n2e_bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/n2e_bench.json; tail -5 gpurun_out/n2e_bench.err
timeout 400 $TR tools/bench_minibatch.py --iters 30 > gpurun_out/n2e_minibatch.json 2> gpurun_out/n2e_minibatch.err; echo "mb2 exit $?"; cat gpurun_out/n2e_minibatch.json; tail -5 gpurun_out/n2e_minibatch.err
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/bench_minibatch.py --iters 30 --cpu-nodes 0 > gpurun_out/n1e_minibatch.json 2> gpurun_out/n1e_minibatch.err; echo "mb1 exit $?"; cat gpurun_out/n1e_minibatch.json; tail -5 gpurun_out/n1e_minibatch.err
