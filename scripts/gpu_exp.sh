#!/bin/bash
mkdir -p gpurun_out
echo "== bnn tests"
timeout 300 python -m pytest tests/test_gpu_sgmcmc.py -q -x -k "bnn" --no-header -p no:cacheprovider 2>&1 | tail -5
echo "== bnn bench"
timeout 300 python scripts/bench_bnn.py 2>&1 | tail -3 | cut -c1-200
echo "== ncu bnn"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sghmc_bnn_kernel -s 30 -c 1 -o gpurun_out/prof_bnn -f python scripts/bench_bnn.py > gpurun_out/ncu_bnn.log 2>&1; tail -1 gpurun_out/ncu_bnn.log
