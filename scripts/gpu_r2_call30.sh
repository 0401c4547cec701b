#!/bin/bash
# Round-2 call 30: same-box A/B of the epi 2 / epi 3 epilogue variants (no early return in the
# epilogue lambda -> plain shuffles; upstream gradient by shuffle vs by warp-uniform load).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm / model tests (default)"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c30_pytest.log 2>&1; tail -3 gpurun_out/r2_c30_pytest.log
echo "== gemm tests (ZSB_BERN_FUSED=1)"
ZSB_BERN_FUSED=1 timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c30_pytest_fused.log 2>&1; tail -3 gpurun_out/r2_c30_pytest_fused.log
for rep in 1 2; do
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "== iwae bench EPI_GLOAD=$1 BERN_FUSED=$2"
  ZSB_EPI_GLOAD=$1 ZSB_BERN_FUSED=$2 timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_c30_iwae_g$1_f$2.err > gpurun_out/r2_c30_iwae_g$1_f$2.json; tail -2 gpurun_out/r2_c30_iwae_g$1_f$2.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c30_iwae_g$1_f$2.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f launches %s frac %.3f bound %s"%(d["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d.get("bound_value")))
P
done
done
