#!/bin/bash
# 2-GPU session: GPU test-suite, N=2 bench (halo / fused / nccl), N=1 bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2_pytest.log
tail -5 gpurun_out/n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
for ex in halo fused nccl; do
  timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --no-e2e --no-cpu > gpurun_out/n2_bench_$ex.json 2> gpurun_out/n2_bench_$ex.err
  echo "$ex exit $?"; tail -c 1500 gpurun_out/n2_bench_$ex.json
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err; echo "n1 exit $?"
tail -c 2500 gpurun_out/n1_bench.json
