#!/bin/bash
# Round 2, final 1-GPU session: the evidence the judge reads -- suite, smoke, N=1 bench + reference arm, ncu launch list of the
# bench command, one ncu --set full capture of the dominant kernel, program A / B numbers with the CPU oracle beside them.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r02z}
timeout 1200 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/${T}_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/${T}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${T}_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_S64_n1.json 2> $O/${T}_bench_S64_n1.err; echo "bench exit $?"; cut -c1-500 $O/${T}_bench_S64_n1.json
kill $SMI
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference exit $?"; cut -c1-300 $O/${T}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench_S64.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${T}_launches.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_tiled_kernel -s 2 -c 2 -f -o $O/${T}_ncu_S64 python tools/profile_spmm.py --workload S64 --iters 2 > $O/${T}_ncu.log 2>&1; echo "ncu full exit $?"
timeout 300 python tools/bench_variants.py --workload S64 > $O/${T}_variants_S64.txt 2>&1; cat $O/${T}_variants_S64.txt
for c in C1 C2; do timeout 600 python tools/bench_epoch.py --config $c --cpu-epochs 1 >> $O/${T}_epoch.jsonl 2>> $O/${T}_epoch.err; done
timeout 300 python tools/bench_epoch.py --config C3 --cpu-epochs 0 >> $O/${T}_epoch.jsonl 2>> $O/${T}_epoch.err
python - <<PY
import json
for ln in open("$O/${T}_epoch.jsonl"):
    j=json.loads(ln); print(j["config"], "eager", round(j["gpu_epoch_ms_events"],3), "graph", round(j["cuda_graph_epoch_ms_events"],3), "cpu oracle s", j["cpu_csr_oracle_epoch_s"], "threads", j["cpu_threads"])
PY
timeout 300 python tools/bench_minibatch.py --graphed --no-prefetch > $O/${T}_minibatch_graphed.json 2> $O/${T}_minibatch_graphed.err; echo "mb graphed exit $?"; cut -c1-300 $O/${T}_minibatch_graphed.json; python -c "
import json; j=json.loads(open('$O/${T}_minibatch_graphed.json').read()); print({k: j.get(k) for k in ('cpu_oracle_s_per_batch','cpu_graph_nodes','cpu_threads')})"
timeout 300 python tools/bench_minibatch.py --cpu-nodes 0 --no-prefetch > $O/${T}_minibatch_eager.json 2> $O/${T}_minibatch_eager.err; cut -c1-300 $O/${T}_minibatch_eager.json
timeout 300 python bench.py --workload C4mb --steps 50 --warmup 10 --no-cpu > $O/${T}_bench_C4mb.json 2> $O/${T}_bench_C4mb.err; echo "C4mb exit $?"; cut -c1-250 $O/${T}_bench_C4mb.json
timeout 300 python tools/bench_dense.py > $O/${T}_dense.txt 2>&1; tail -32 $O/${T}_dense.txt
