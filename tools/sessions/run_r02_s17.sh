#!/bin/bash
# round 2, session 17b: rows staged through shared memory by TMA (tile::gather4 / per-row bulk copy) vs the shipped kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tma_row" > gpurun_out/r02w_pytest_tma.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r02w_pytest_tma.log
timeout 300 python tools/bench_variants.py --tma-ab > gpurun_out/r02w_tma_rows_ab.txt 2> gpurun_out/r02w_tma_rows_ab.err; echo "ab exit $?"
cat gpurun_out/r02w_tma_rows_ab.txt; tail -3 gpurun_out/r02w_tma_rows_ab.err
