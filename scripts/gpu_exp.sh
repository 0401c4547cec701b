#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_new.log
timeout 400 python scripts/bench_linear.py 2>&1 | tail -1 | tee gpurun_out/bench_linear.log
echo "== iwae"
timeout 400 python scripts/bench_iwae.py 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench_iwae.log
