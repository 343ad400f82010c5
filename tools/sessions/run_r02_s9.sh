#!/bin/bash
# Round 2, GPU session 9 (1 GPU): hybrid register + cp.async shared-memory staging of the gathered rows (more bytes in flight
# than the register budget allows): parity of the variant libraries, then S64 / C4 / uniform timings against the shipped kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in "" ab_smem2 ab_smem4; do
  tag=${lib:-shipped}
  export GGAD_B200_LIB=${lib:+$PWD/$lib.so}
  [ -z "$lib" ] && unset GGAD_B200_LIB
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "widths or epilogues or tiled or autograd_large or exchange" > $O/r02j_pytest_$tag.log 2>&1; echo "$tag pytest exit $?"; tail -2 $O/r02j_pytest_$tag.log
  timeout 300 python tools/bench_variants.py --workload S64 --no-chase > $O/r02j_variants_S64_$tag.txt 2>&1; grep -i "plain\|GCN\|no z\|push variant, 42" $O/r02j_variants_S64_$tag.txt | sed "s/^/$tag /"
  for wl in S64 C4; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-cpu > $O/r02j_bench_${wl}_$tag.json 2> $O/r02j_bench_${wl}_$tag.err
    python - <<PY
import json
j=json.loads(open("$O/r02j_bench_${wl}_$tag.json").read().strip().splitlines()[-1])
print("$tag $wl", "fwd", round(j["segments_ms"]["fwd_compute"],3), "bwd", round(j["segments_ms"]["bwd_compute"],3), "GE/s", round(j["value"]/1e9,2), "verify", j["verified_rows"]["max_err_over_bound"])
PY
  done
  timeout 300 python bench.py --workload S64 --rmat 0.25,0.25,0.25 --steps 10 --warmup 3 --no-e2e --no-cpu > $O/r02j_bench_uniform_$tag.json 2>/dev/null
  python - <<PY
import json
j=json.loads(open("$O/r02j_bench_uniform_$tag.json").read().strip().splitlines()[-1])
print("$tag uniform", "fwd", round(j["segments_ms"]["fwd_compute"],3), "bwd", round(j["segments_ms"]["bwd_compute"],3))
PY
done
