#!/bin/bash
mkdir -p gpurun_out
echo "== iwae bench"
timeout 600 python scripts/bench_iwae.py 2>&1 | tail -3 | tee gpurun_out/bench_iwae.json | cut -c1-900
