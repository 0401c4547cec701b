#!/bin/bash
# Round-2 call 5: P-tile multicast variant of the resident kernel (clusters of 4).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== MC tests"
ZSB_RES_MC=1 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "resident_kernel or (trajectory_kernels and 5) or (golden and 5)" > gpurun_out/r2_mc_tests.log 2>&1; tail -12 gpurun_out/r2_mc_tests.log
echo "== golden 64 impl 0 (tolerance fix)"
timeout 300 python -m pytest tests/test_gpu_hmc.py -m gpu -q --no-header -p no:cacheprovider -k "golden_dense64" 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --dense-impl 5"
for mc in 1 0 1 0; do
  echo "== sustained res MC=$mc"
  ZSB_RES_MC=$mc timeout 300 $B 2> gpurun_out/r2_mc$mc.err > gpurun_out/r2_mc$mc.json; tail -2 gpurun_out/r2_mc$mc.err; python scripts/show_bench.py gpurun_out/r2_mc$mc.json | head -3
done
echo "== short full-clock runs"
for mc in 1 0; do
  ZSB_RES_MC=$mc timeout 300 python bench.py --steps 3 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5 2> gpurun_out/r2_mcs$mc.err > gpurun_out/r2_mcs$mc.json; tail -2 gpurun_out/r2_mcs$mc.err; python scripts/show_bench.py gpurun_out/r2_mcs$mc.json | head -3
done
