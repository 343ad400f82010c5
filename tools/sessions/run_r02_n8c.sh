#!/bin/bash
# Round 2, third 8-GPU session (short): program B data parallel with the two-graph step on 8 GPUs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 200 > $O/r02t_minibatch_graphed_n8.json 2> $O/r02t_minibatch_graphed_n8.err; echo "mb graphed n8 exit $?"; cut -c1-400 $O/r02t_minibatch_graphed_n8.json; tail -2 $O/r02t_minibatch_graphed_n8.err
