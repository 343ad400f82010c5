#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2g_pytest.log
tail -4 gpurun_out/n2g_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR tools/bench_minibatch.py --iters 40 --diag > gpurun_out/n2g_minibatch_diag.json 2> gpurun_out/n2g_minibatch_diag.err; echo "mb2 diag exit $?"; cat gpurun_out/n2g_minibatch_diag.json; tail -3 gpurun_out/n2g_minibatch_diag.err
timeout 400 $TR tools/bench_minibatch.py --iters 60 > gpurun_out/n2g_minibatch.json 2> gpurun_out/n2g_minibatch.err; echo "mb2 exit $?"; cat gpurun_out/n2g_minibatch.json
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/bench_minibatch.py --iters 60 --diag > gpurun_out/n1g_minibatch.json 2> gpurun_out/n1g_minibatch.err; echo "mb1 exit $?"; cat gpurun_out/n1g_minibatch.json
