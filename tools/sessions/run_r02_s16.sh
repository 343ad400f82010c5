#!/bin/bash
# Round 2, GPU session 16 (2 GPUs): two-graph data-parallel mini-batch step -- parity (multi-GPU DP test + single-GPU graphed
# test), batches/s at N=2 against the eager data-parallel step.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_modules.py -m gpu -q -k "data_parallel or graphed" > $O/r02s_pytest.log 2>&1; echo "pytest exit $?"; grep -n "AssertionError\|passed\|failed\|^E " $O/r02s_pytest.log | cut -c1-500 | head
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 200 > $O/r02s_minibatch_graphed_n2.json 2> $O/r02s_minibatch_graphed_n2.err; echo "mb graphed n2 exit $?"; cut -c1-330 $O/r02s_minibatch_graphed_n2.json; tail -2 $O/r02s_minibatch_graphed_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch --iters 200 > $O/r02s_minibatch_graphed_n1.json 2> /dev/null; cut -c1-330 $O/r02s_minibatch_graphed_n1.json
