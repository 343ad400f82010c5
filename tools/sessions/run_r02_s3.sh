#!/bin/bash
# Round 2, GPU session 3 (2 GPUs): whole parity suite incl. the multi-GPU tests (halo, chase, hybrid multicast, oracle
# check), then the S64 bench at N=2 with the in-kernel halo push vs the chase exchange.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest exit $?"; tail -12 $O/r02c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
for ex in halo chase; do
  timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex --no-cpu > $O/r02c_n2_$ex.json 2> $O/r02c_n2_$ex.err; echo "bench $ex exit $?"
  python - <<PY
import json
j=json.loads(open("$O/r02c_n2_$ex.json").read().strip().splitlines()[-1])
print("$ex", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", j["segments_ms"]["per_rank"], "e2e", round(j["e2e"]["ms_per_step"],2), j["verified_rows"])
PY
done
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange chase --chase-ctas 96 --no-cpu --no-e2e > $O/r02c_n2_chase96.json 2> $O/r02c_n2_chase96.err; echo "bench chase96 exit $?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange chase --mc-min 1 --no-cpu --no-e2e > $O/r02c_n2_chase_mc.json 2> $O/r02c_n2_chase_mc.err; echo "bench chase mc exit $?"
for f in chase96 chase_mc; do python - <<PY
import json
j=json.loads(open("$O/r02c_n2_$f.json").read().strip().splitlines()[-1])
print("$f", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", j["segments_ms"]["per_rank"])
PY
done
timeout 300 python tools/bench_dense.py > $O/r02c_dense.txt 2>&1; tail -8 $O/r02c_dense.txt
for c in C1 C2 C3; do CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_epoch.py --config $c --cpu-epochs 0 >> $O/r02c_epoch.jsonl 2>> $O/r02c_epoch.err; done; echo "epoch exit $?"; tail -c 1200 $O/r02c_epoch.jsonl
