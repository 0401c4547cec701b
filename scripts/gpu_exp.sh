#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_distributions.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_new.log
