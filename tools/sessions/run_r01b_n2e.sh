#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/n2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n2f_pytest.log
tail -6 gpurun_out/n2f_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2f_bench.json 2> gpurun_out/n2f_bench.err; echo "bench exit $?"; python - <<PY
import json
j=json.loads(open("gpurun_out/n2f_bench.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"]/1e9, j["e2e"], j["clocks"])
PY
tail -3 gpurun_out/n2f_bench.err
timeout 400 $TR tools/bench_minibatch.py --iters 40 --diag > gpurun_out/n2f_minibatch_diag.json 2> gpurun_out/n2f_minibatch_diag.err; echo "mb2 exit $?"; cat gpurun_out/n2f_minibatch_diag.json; tail -3 gpurun_out/n2f_minibatch_diag.err
OMP_NUM_THREADS=8 timeout 400 $TR tools/bench_minibatch.py --iters 40 > gpurun_out/n2f_minibatch_omp8.json 2> gpurun_out/n2f_minibatch_omp8.err; echo "mb2 omp8 exit $?"; cat gpurun_out/n2f_minibatch_omp8.json
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/bench_minibatch.py --iters 40 --cpu-nodes 0 --diag > gpurun_out/n1f_minibatch_diag.json 2> gpurun_out/n1f_minibatch_diag.err; echo "mb1 exit $?"; cat gpurun_out/n1f_minibatch_diag.json
CUDA_VISIBLE_DEVICES=0 OMP_NUM_THREADS=1 timeout 400 python tools/bench_minibatch.py --iters 40 --cpu-nodes 0 > gpurun_out/n1f_minibatch_omp1.json 2> gpurun_out/n1f_minibatch_omp1.err; echo "mb1 omp1 exit $?"; cat gpurun_out/n1f_minibatch_omp1.json
for wl in C4; do
 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload $wl --steps 10 --no-e2e --no-cpu > gpurun_out/n1f_bench_$wl.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/n1f_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl n1', j['ms_per_step'], j['value']/1e9, j['segments_ms']['per_rank'])"
 timeout 300 $TR bench.py --gpus 2 --workload $wl --steps 10 --no-e2e --no-cpu > gpurun_out/n2f_bench_$wl.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/n2f_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl n2', j['ms_per_step'], j['value']/1e9, j['segments_ms']['per_rank'], j['config']['halo_rows_sent_frac'])"
done
