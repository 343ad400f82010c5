#!/bin/bash
# round 2, session 21b: model_ocgnn encoder + tam.py affinity drop-ins against the reference-generated goldens (one short call)
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_modules.py -x -q -m gpu -k "encoder_model" > gpurun_out/r02zz_pytest_enc.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02zz_pytest_enc.log | cut -c1-300
