#!/bin/bash
# single-GPU profile session: launch list of the bench command, one ncu --set full capture of the forward and
# backward gather kernels (+ the push variant), final N=1 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r01b_bench_S64_n1.json 2> gpurun_out/r01b_bench_S64_n1.err; echo "bench exit $?"
tail -c 2600 gpurun_out/r01b_bench_S64_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_bench_S64.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01b_launches.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_tiled_kernel -s 2 -c 2 -f -o gpurun_out/r01b_ncu_S64 python tools/profile_spmm.py --workload S64 --iters 2 > gpurun_out/r01b_ncu.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/bench_variants.py --workload S64 > gpurun_out/r01b_variants_S64.txt 2>&1; cat gpurun_out/r01b_variants_S64.txt
for c in C1 C2 C3; do timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> gpurun_out/r01b_epoch.jsonl 2>> gpurun_out/r01b_epoch.err; done; echo "epoch exit $?"; tail -c 1800 gpurun_out/r01b_epoch.jsonl
