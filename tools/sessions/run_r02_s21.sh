#!/bin/bash
# round 2, session 21: tam.py affinity drop-ins against the reference-generated goldens (one short call)
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_modules.py -x -q -m gpu -k "tam_affinity or affinity_all_rows" > gpurun_out/r02zz_pytest_tam.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02zz_pytest_tam.log | cut -c1-300
