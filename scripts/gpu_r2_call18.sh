#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sgmcmc.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_call18_tests.log
cat gpurun_out/r2_call18_tests.log
timeout 300 python scripts/bench_bnn.py > gpurun_out/r2_call18_bnn.jsonl 2> gpurun_out/r2_call18_bnn.err; cat gpurun_out/r2_call18_bnn.jsonl; tail -3 gpurun_out/r2_call18_bnn.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sghmc_bnn -s 5 -c 1 -o gpurun_out/r2_call18_bnn python scripts/bench_bnn.py > gpurun_out/r2_call18_ncu.log 2>&1; tail -3 gpurun_out/r2_call18_ncu.log
