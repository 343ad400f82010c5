#!/bin/bash
# Round 2, 8-GPU session: S64 x 8 = the C5 graph (50 M nodes / 1 B edges): layer pass with the in-kernel halo push and with
# the chase exchange + hybrid multicast; C5 as BASELINE.json words it (two-layer mean-SAGE mini-batches, feature table
# sharded, one NCCL all-reduce per layer) next to the replicated-table variant; program B data parallel with the input pipeline.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02g_n${N}_halo.json 2> $O/r02g_n${N}_halo.err; echo "bench halo exit $?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --exchange chase --mc-min 4 --no-cpu --no-e2e > $O/r02g_n${N}_chase_mc4.json 2> $O/r02g_n${N}_chase_mc4.err; echo "bench chase mc4 exit $?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --exchange chase --no-cpu --no-e2e > $O/r02g_n${N}_chase.json 2> $O/r02g_n${N}_chase.err; echo "bench chase exit $?"
for f in halo chase_mc4 chase; do python - <<PY
import json
try:
    j=json.loads(open("$O/r02g_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"],3), "ms", round(j["value"]/1e9,2), "GE/s", [r[:3:2] for r in j["segments_ms"]["per_rank"]], "e2e", j.get("e2e") and round(j["e2e"]["ms_per_step"],1), j["verified_rows"])
except Exception as e: print("$f", "failed", e)
PY
done
timeout 900 $TR tools/bench_sharded_sage.py --mode both --iters 20 --warm 5 > $O/r02g_sharded_sage_n${N}.json 2> $O/r02g_sharded_sage_n${N}.err; echo "sharded sage exit $?"; tail -c 1600 $O/r02g_sharded_sage_n${N}.json; tail -3 $O/r02g_sharded_sage_n${N}.err
timeout 600 $TR tools/bench_minibatch.py --cpu-nodes 0 > $O/r02g_minibatch_n${N}.json 2> $O/r02g_minibatch_n${N}.err; echo "minibatch exit $?"; cut -c1-700 $O/r02g_minibatch_n${N}.json; tail -3 $O/r02g_minibatch_n${N}.err
