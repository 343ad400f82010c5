#!/bin/bash
# round 2, session 20: two request paths at once (TMA gather4 warps + register-direct warps in every CTA)
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tma_row or mixed_request" > $O/r02y_pytest_mix.log 2>&1; echo "pytest exit $?"
tail -4 $O/r02y_pytest_mix.log
timeout 300 python tools/bench_variants.py --tma-ab > $O/r02y_mixed_paths_ab.txt 2> $O/r02y_mixed_paths_ab.err; echo "ab exit $?"
cat $O/r02y_mixed_paths_ab.txt; tail -3 $O/r02y_mixed_paths_ab.err
