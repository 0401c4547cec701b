#!/bin/bash
# Round-2 call 29: epi 2 / epi 3 epilogues without shuffles (uniform loads of the upstream gradient,
# packed fp16 conversions); fused Bernoulli backward opt-in vs default.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm / model tests (default)"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c29_pytest.log 2>&1; tail -6 gpurun_out/r2_c29_pytest.log
echo "== gemm / model tests (ZSB_BERN_FUSED=1)"
ZSB_BERN_FUSED=1 timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c29_pytest_fused.log 2>&1; tail -6 gpurun_out/r2_c29_pytest_fused.log
for t in 0 1 0 1; do
  echo "== iwae bench BERN_FUSED=$t"
  ZSB_BERN_FUSED=$t timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_c29_iwae_f$t.err > gpurun_out/r2_c29_iwae_f$t.json; tail -2 gpurun_out/r2_c29_iwae_f$t.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c29_iwae_f$t.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.4e launches %s frac %.3f mma %.3f bound %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["frac"],d["roofline"]["mma_issued_frac_of_peak"],d.get("bound_value")))
P
done
