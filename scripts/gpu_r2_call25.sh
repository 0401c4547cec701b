#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_sgmcmc.py -q -m gpu -k "reference_run" 2>&1 | tail -30 > gpurun_out/r2_call25_ref.log
cat gpurun_out/r2_call25_ref.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_call25_tests.log
cat gpurun_out/r2_call25_tests.log
