#!/bin/bash
# Round-2 call 26: 128-bit momentum / kinetic / select_planes kernels, parallel mass_stats stage 2.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider -x > gpurun_out/r2_c26_pytest.log 2>&1; tail -6 gpurun_out/r2_c26_pytest.log
echo "== kernel microbench"
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_c26_bench_kernels.jsonl 2> gpurun_out/r2_c26_bench_kernels.err; tail -3 gpurun_out/r2_c26_bench_kernels.err; cut -c1-200 gpurun_out/r2_c26_bench_kernels.jsonl
echo "== headline bench"
timeout 600 python bench.py --steps 10 --warmup 5 2> gpurun_out/r2_c26_bench.err > gpurun_out/r2_c26_bench.json; tail -2 gpurun_out/r2_c26_bench.err; python scripts/show_bench.py gpurun_out/r2_c26_bench.json | head -12
