#!/bin/bash
# Round 2, second 8-GPU session: hybrid multicast inside the in-kernel halo push (rows needed by >= k peers go once through
# the NVSwitch multicast address), against the plain halo push; multi-GPU tests at world 4.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k sharded_layer > $O/r02n_pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -n "AssertionError\|passed\|failed" $O/r02n_pytest_multi.log | cut -c1-800
for mc in 0 2 4 6; do
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --mc-min $mc --no-cpu --no-e2e > $O/r02n_n${N}_halo_mc$mc.json 2> $O/r02n_n${N}_halo_mc$mc.err; echo "bench mc$mc exit $?"
  python - <<PY
import json
try:
    j=json.loads(open("$O/r02n_n${N}_halo_mc$mc.json").read().strip().splitlines()[-1])
    print("mc_min=$mc", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s fwd", [r[0] for r in j["segments_ms"]["per_rank"]], j["verified_rows"]["halo_rows_bit_exact"], j["verified_rows"]["max_err_over_bound"])
except Exception as e: print("mc_min=$mc failed", e)
PY
done
