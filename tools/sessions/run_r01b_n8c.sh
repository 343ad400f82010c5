#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 tools/bench_minibatch.py --iters 100 --diag > gpurun_out/n8c_minibatch.json 2> gpurun_out/n8c_minibatch.err; echo "mb8 exit $?"; cat gpurun_out/n8c_minibatch.json; tail -2 gpurun_out/n8c_minibatch.err
