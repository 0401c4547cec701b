#!/bin/bash
mkdir -p gpurun_out
echo "== bnn tests"
timeout 300 python -m pytest tests/test_gpu_sgmcmc.py -q -x -k "bnn" --no-header -p no:cacheprovider 2>&1 | tail -25
echo "== bnn bench"
timeout 300 python scripts/bench_bnn.py 2>&1 | tail -6
