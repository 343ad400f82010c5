#!/bin/bash
# Round 2, last 2-GPU check on the final library: multi-GPU parity tests + the default bench line at N=2.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02zzz_pytest_multi.log 2>&1; echo "pytest exit $?"; tail -2 $O/r02zzz_pytest_multi.log | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02zzz_scale_n2.json 2> $O/r02zzz_scale_n2.err; echo "bench n2 exit $?"
python - <<PY
import json
j=json.loads(open("$O/r02zzz_scale_n2.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e9,2), "GE/s", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],2), j["verified_rows"])
PY
