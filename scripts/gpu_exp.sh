#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
timeout 300 python -m pytest tests/test_gpu_hmc.py -m gpu -q -x --no-header -p no:cacheprovider -k "single_pass and (1000 or 130 or 300)" 2>&1 | tail -15 | tee gpurun_out/pytest_impl3.log
for impl in 3 2; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-adapt --burnin 2 --no-e2e --no-cpu-baseline --dense-impl $impl 2>gpurun_out/b_impl$impl.err > gpurun_out/b_impl$impl.json; python scripts/show_bench.py gpurun_out/b_impl$impl.json | head -3; tail -2 gpurun_out/b_impl$impl.err
done
