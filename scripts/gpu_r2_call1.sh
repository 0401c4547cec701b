#!/bin/bash
# Round-2 call 1: first execution of the trajectory-fused dense kernel (impl 4).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== trajectory tests"
ZSB_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "trajectory" 2>&1 | tail -30 | tee gpurun_out/r2_traj_test.log
echo "== bench impl 4"
timeout 600 python bench.py --steps 5 --warmup 3 --dense-impl 4 --no-cpu-baseline 2> gpurun_out/r2_b_impl4.err > gpurun_out/r2_b_impl4.json; tail -5 gpurun_out/r2_b_impl4.err; python scripts/show_bench.py gpurun_out/r2_b_impl4.json
echo "== bench impl 2"
timeout 600 python bench.py --steps 5 --warmup 3 --dense-impl 2 --no-cpu-baseline 2> gpurun_out/r2_b_impl2.err > gpurun_out/r2_b_impl2.json; tail -5 gpurun_out/r2_b_impl2.err; python scripts/show_bench.py gpurun_out/r2_b_impl2.json
