#!/bin/bash
# reference-run fixtures replayed on the CUDA paths, then the full GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_sgmcmc.py tests/test_gpu_models.py -q -m gpu -k "reference_run or ref_sgmcmc" 2>&1 | tail -25 > gpurun_out/r2_call16_ref.log
cat gpurun_out/r2_call16_ref.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_call16_full.log
cat gpurun_out/r2_call16_full.log
