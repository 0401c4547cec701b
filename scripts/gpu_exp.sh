#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py 2>&1 | tail -30 | tee gpurun_out/bench_kernels.jsonl
