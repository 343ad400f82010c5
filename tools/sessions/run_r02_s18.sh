#!/bin/bash
# round 2, session 18: TMA-staged rows, wider A/B (1 stage, L2 promotion off) + ncu of the 2- and 4-stage gather4 kernels
O=gpurun_out; mkdir -p $O
timeout 300 python tools/bench_variants.py --tma-ab > $O/r02v_tma_rows_ab.txt 2> $O/r02v_tma_rows_ab.err; echo "ab exit $?"
cat $O/r02v_tma_rows_ab.txt; tail -3 $O/r02v_tma_rows_ab.err
for st in 2 4; do
  GGAD_TMA_ROWS=1 GGAD_TMA_STAGES=$st timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_tma_kernel -s 1 -c 1 -f \
    -o $O/r02v_ncu_S64_tma_gather4_s$st python tools/profile_spmm.py --workload S64 --iters 2 > $O/r02v_ncu_s$st.log 2>&1; echo "ncu s$st exit $?"
done
ls -la $O/*.ncu-rep | tail -3
