#!/bin/bash
# round 2, session 19: default bench line with the concurrent host-link probe
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/r02x_bench_S64_n1.json 2> $O/r02x_bench_S64_n1.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$O/r02x_bench_S64_n1.json").read().strip().splitlines()[-1])
print(round(j["value"]/1e9,2), "GE/s", round(j["ms_per_step"],3), "ms  e2e", round(j["e2e"]["ms_per_step"],2), j["e2e"]["host_link_GBs_achieved"], j["e2e"]["host_link"])
PY
for s in 2 4; do timeout 300 python bench.py --e2e-slots $s --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slots $s e2e', round(j['e2e']['ms_per_step'],2), j['e2e']['host_link_GBs_achieved'])"; done
