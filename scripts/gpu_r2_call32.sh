#!/bin/bash
# Round-2 call 32: input gradient from the forward weight planes (A operand MN-major), ticketed
# absmax (no pow2 kernel after a max pass).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm / model / estimator tests"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c32_pytest.log 2>&1; tail -12 gpurun_out/r2_c32_pytest.log
for rep in 1 2 3; do
  echo "== iwae bench"
  timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_c32_iwae.err > gpurun_out/r2_c32_iwae_$rep.json; tail -2 gpurun_out/r2_c32_iwae.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c32_iwae_$rep.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.4e launches %s frac %.3f bound %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["frac"],d.get("bound_value")))
P
done
echo "== iwae bench, round-2 scheme (ZSB_WGRAD_T=1) on the same box"
ZSB_WGRAD_T=1 timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_c32_iwae.err > gpurun_out/r2_c32_iwae_t.json; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c32_iwae_t.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f launches %s"%(d["value"],d["ms_per_step"],d["gpu_launches"]))
P
