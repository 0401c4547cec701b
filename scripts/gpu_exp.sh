#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_distributions.py -m gpu -q --no-header -p no:cacheprovider -k "multinomial_and_onehot" 2>&1 | tail -12 | tee gpurun_out/pytest_new.log
timeout 100 python scripts/bench_iwae.py 2>&1 | tail -1 | cut -c1-700 | tee gpurun_out/bench_iwae2.log
