#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_call22_tests.log
cat gpurun_out/r2_call22_tests.log
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_call22_kernels.jsonl 2> gpurun_out/r2_call22_kernels.err; grep -i "bernoulli\|normal\|mean_exp\|momentum\|mass" gpurun_out/r2_call22_kernels.jsonl; tail -3 gpurun_out/r2_call22_kernels.err
timeout 300 python bench.py --workload iwae --steps 10 --warmup 3 > gpurun_out/r2_call22_iwae.json 2> gpurun_out/r2_call22_iwae.err; cut -c1-400 gpurun_out/r2_call22_iwae.json; tail -2 gpurun_out/r2_call22_iwae.err
