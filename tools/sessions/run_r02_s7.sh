#!/bin/bash
# Round 2, GPU session 7 (2 GPUs): suite (multi-GPU oracle numbers, graphed mini-batch step), program B with the CUDA-graph
# tail at N=1 and N=2, GEMM dispatch table, C5-shaped sharded SAGE with entry-balanced column ranges at N=2.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02h_pytest.log 2>&1; echo "pytest exit $?"; tail -30 $O/r02h_pytest.log | cut -c1-1500
export CUDA_VISIBLE_DEVICES=0
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed > $O/r02h_minibatch_graphed.json 2> $O/r02h_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-700 $O/r02h_minibatch_graphed.json; tail -3 $O/r02h_minibatch_graphed.err
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --graphed --no-prefetch > $O/r02h_minibatch_graphed_nopf.json 2> $O/r02h_minibatch_graphed_nopf.err; echo "mb graphed no-prefetch exit $?"; cut -c1-500 $O/r02h_minibatch_graphed_nopf.json
timeout 300 python tools/bench_dense.py > $O/r02h_dense.txt 2>&1; cat $O/r02h_dense.txt
unset CUDA_VISIBLE_DEVICES
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 --graphed > $O/r02h_minibatch_graphed_n2.json 2> $O/r02h_minibatch_graphed_n2.err; echo "mb graphed n2 exit $?"; cut -c1-500 $O/r02h_minibatch_graphed_n2.json; tail -3 $O/r02h_minibatch_graphed_n2.err
timeout 900 $TR tools/bench_sharded_sage.py --mode sharded --iters 20 --warm 5 > $O/r02h_sharded_sage_n2.json 2> $O/r02h_sharded_sage_n2.err; echo "sharded sage exit $?"; tail -c 900 $O/r02h_sharded_sage_n2.json; tail -3 $O/r02h_sharded_sage_n2.err
