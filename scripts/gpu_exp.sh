#!/bin/bash
mkdir -p gpurun_out
echo "== model tests"
timeout 600 python -m pytest tests/test_gpu_models.py -q --no-header -p no:cacheprovider 2>&1 | tail -25
