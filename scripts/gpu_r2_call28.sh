#!/bin/bash
# Round-2 call 28: epi 3 (Bernoulli d/dlogits emitted as operand planes from the GEMM epilogue).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== gemm / model tests"
timeout 600 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_estimators.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_c28_pytest.log 2>&1; tail -25 gpurun_out/r2_c28_pytest.log
for t in 1 0 1 0; do
  echo "== iwae bench BERN_UNFUSED=$t"
  ZSB_BERN_UNFUSED=$t timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_c28_iwae_u$t.err > gpurun_out/r2_c28_iwae_u$t.json; tail -2 gpurun_out/r2_c28_iwae_u$t.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2_c28_iwae_u$t.json").read().strip().splitlines()[-1])
print("value %.4e ms %.3f e2e %.4e launches %s frac %.3f mma %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"],d["roofline"]["frac"],d["roofline"]["mma_issued_frac_of_peak"]))
P
done
echo "== launch list of the iwae step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_c28_iwae_launches.csv python bench.py --workload iwae --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-cuda-graph > gpurun_out/r2_c28_ncu.log 2>&1; tail -2 gpurun_out/r2_c28_ncu.log
python scripts/summarize_launches.py gpurun_out/r2_c28_iwae_launches.csv 2>/dev/null | head -30
