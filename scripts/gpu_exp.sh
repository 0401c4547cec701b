#!/bin/bash
# validation of the new estimator / distribution / diagnostics kernels + dense-kernel experiments
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
echo "== new tests"
timeout 900 python -m pytest tests/test_gpu_estimators.py tests/test_gpu_distributions.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_new.log
for dbg in 1 2; do
echo "== dbg $dbg"
ZSB_TC_DBG=$dbg timeout 400 python bench.py --steps 10 --warmup 3 --no-adapt --burnin 2 --no-e2e --no-cpu-baseline 2>gpurun_out/b_dbg$dbg.err > gpurun_out/b_dbg$dbg.json; python scripts/show_bench.py gpurun_out/b_dbg$dbg.json | head -3; tail -2 gpurun_out/b_dbg$dbg.err
done
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err > gpurun_out/bench.json; python scripts/show_bench.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "== iwae (cublas fp32 emulation)"
CUBLAS_EMULATE_SINGLE_PRECISION=1 timeout 300 python scripts/bench_iwae.py 2>&1 | tail -2 | cut -c1-900
