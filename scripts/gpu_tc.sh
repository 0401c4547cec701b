#!/bin/bash
# TC kernel bring-up: targeted tests first (short timeout: the kernel traps after ~2 s on a stuck
# barrier), then the full GPU suite and both bench variants.
mkdir -p gpurun_out
echo "== tc single-pass tests"
timeout 300 python -m pytest tests/test_gpu_hmc.py -q -x -k "single_pass" --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/tc_single.log
echo "== compute-sanitizer (small)"
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_hmc.py -q -x -k "single_pass and 24-32" --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/tc_sanitizer.log
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench impl1"
timeout 600 python bench.py --steps 5 --warmup 3 --burnin 14 --dense-impl 1 2> gpurun_out/bench1.err | tee gpurun_out/bench_impl1.json; tail -5 gpurun_out/bench1.err
echo "== bench impl0"
timeout 600 python bench.py --steps 3 --warmup 3 --burnin 14 --dense-impl 0 --no-cpu-baseline 2> gpurun_out/bench0.err | tee gpurun_out/bench_impl0.json; tail -3 gpurun_out/bench0.err
