#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_distributions.py -m gpu -q --no-header -p no:cacheprovider -k "plugin_distribution or normal" 2>&1 | tail -15 | tee gpurun_out/pytest_new.log
