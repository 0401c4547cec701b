#!/bin/bash
# Round-2 call 32: opt-in paths -- input gradient from the forward weight planes (A operand
# MN-major, ZSB_DGRAD_MN=1) and ticketed absmax (ZSB_ABSMAX_TICKET=1) -- vs the default build.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm / model / estimator tests (default)"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c32_pytest.log 2>&1; tail -4 gpurun_out/r2_c32_pytest.log
echo "== the same with ZSB_DGRAD_MN=1 ZSB_ABSMAX_TICKET=1"
ZSB_DGRAD_MN=1 ZSB_ABSMAX_TICKET=1 timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c32_pytest_optin.log 2>&1; tail -12 gpurun_out/r2_c32_pytest_optin.log
for cfg in "0 0" "1 1" "0 0" "1 1" "1 0"; do
  set -- $cfg
  echo "== iwae bench DGRAD_MN=$1 ABSMAX_TICKET=$2"
  ZSB_DGRAD_MN=$1 ZSB_ABSMAX_TICKET=$2 timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_c32_iwae.err > gpurun_out/r2_c32_iwae_d$1_t$2.json; tail -2 gpurun_out/r2_c32_iwae.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c32_iwae_d$1_t$2.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f launches %s frac %.3f bound %s"%(d["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d.get("bound_value")))
P
done
