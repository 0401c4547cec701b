#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_distributions.py -m gpu -q --no-header -p no:cacheprovider -k "concrete_family or multinomial_and_onehot" 2>&1 | tail -25 | tee gpurun_out/pytest_new.log
