#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_estimators.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2_call24_est.log
cat gpurun_out/r2_call24_est.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_call24_tests.log
cat gpurun_out/r2_call24_tests.log
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_call24_kernels.jsonl 2> gpurun_out/r2_call24_kernels.err; grep -i "bernoulli\|normal\|mass\|kinetic" gpurun_out/r2_call24_kernels.jsonl; tail -3 gpurun_out/r2_call24_kernels.err
