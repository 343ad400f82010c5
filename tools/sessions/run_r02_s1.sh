#!/bin/bash
# Round 2, GPU session 1 (1 GPU): parity suite incl. the new config-shape / chase / golden-big tests, N=1 bench with
# the in-bench verification, A/B of the epilogue kinds and of the exchange variants on S64 (local stand-in peers),
# ncu --set full of the variants programs A/B actually run, reference arm on the full S64 config.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench_S64_n1.json 2> $O/r02a_bench_S64_n1.err; echo "bench exit $?"; tail -c 1500 $O/r02a_bench_S64_n1.json; tail -3 $O/r02a_bench_S64_n1.err
for tag in shipped full deferred; do
  env=""
  [ $tag = full ] && env="GGAD_FORCE_FULL_EPI=1"
  [ $tag = deferred ] && env="GGAD_EPI_DEFERRED=1"
  env $env timeout 300 python tools/bench_variants.py --workload S64 > $O/r02a_variants_S64_$tag.txt 2>&1; echo "variants $tag exit $?"
  cat $O/r02a_variants_S64_$tag.txt
done
for lib in ab_light_rolled ab_push_group; do
  GGAD_B200_LIB=$PWD/$lib.so timeout 300 python tools/bench_variants.py --workload S64 > $O/r02a_variants_S64_$lib.txt 2>&1; echo "variants $lib exit $?"
  grep -i "GCN\|no z\|push\|plain" $O/r02a_variants_S64_$lib.txt
done
GGAD_EPI_DEFERRED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -m gpu -x -q > $O/r02a_pytest_deferred.log 2>&1; echo "pytest deferred exit $?"; tail -3 $O/r02a_pytest_deferred.log
GGAD_B200_LIB=$PWD/ab_push_group.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "exchange" > $O/r02a_pytest_pushgroup.log 2>&1; echo "pytest push_group exit $?"; tail -3 $O/r02a_pytest_pushgroup.log
for v in gcn_layer col_scale push; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_tiled_kernel -s 1 -c 1 -f -o $O/r02a_ncu_S64_$v python tools/profile_spmm.py --workload S64 --iters 2 --variant $v > $O/r02a_ncu_$v.log 2>&1; echo "ncu $v exit $?"
done
GGAD_EPI_DEFERRED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_tiled_kernel -s 1 -c 1 -f -o $O/r02a_ncu_S64_gcn_layer_deferred python tools/profile_spmm.py --workload S64 --iters 2 --variant gcn_layer > $O/r02a_ncu_gcn_deferred.log 2>&1; echo "ncu deferred exit $?"
ls -la $O/*.ncu-rep
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02a_bench_reference.json 2> $O/r02a_bench_reference.err; echo "reference exit $?"; cat $O/r02a_bench_reference.json | cut -c1-900
free -g | head -2; nproc
