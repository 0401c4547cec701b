#!/bin/bash
# Round-2 call 27: weight gradient from row-major planes (MN-major tcgen05 operands).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm tests"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c27_pytest.log 2>&1; tail -25 gpurun_out/r2_c27_pytest.log
for t in 1 0 1 0; do
  echo "== iwae bench WGRAD_T=$t"
  ZSB_WGRAD_T=$t timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_c27_iwae_t$t.err > gpurun_out/r2_c27_iwae_t$t.json; tail -2 gpurun_out/r2_c27_iwae_t$t.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c27_iwae_t$t.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.4e launches %s frac %.3f mma %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["frac"],d["roofline"]["mma_issued_frac_of_peak"]))
P
done
