#!/bin/bash
# Round-2 call 11: K8 dense layers with fused memory passes; IWAE eager + graph; checkpoint tests.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== K8 / models / state_dict tests"
timeout 900 python -m pytest tests/test_gpu_gemm_logjoint.py tests/test_gpu_models.py tests/test_gpu_sgmcmc.py tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "gemm or linear or iwae or vae or state_dict or bnn or ais" 2>&1 | tail -15
echo "== iwae eager"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_iwae_eager2.err > gpurun_out/r2_iwae_eager2.json; tail -3 gpurun_out/r2_iwae_eager2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_eager2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['bound_value'])"
echo "== iwae graph"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline --cuda-graph 2> gpurun_out/r2_iwae_graph2.err > gpurun_out/r2_iwae_graph2.json; tail -5 gpurun_out/r2_iwae_graph2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_graph2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['bound_value'])"
echo "== iwae launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_iwae_launches2.csv python bench.py --workload iwae --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_iwae_ncu2.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_iwae_launches2.csv 2>/dev/null | head -16
echo "== scripts/bench_iwae.py (gradient accuracy vs fp32 matmul)"
timeout 600 python scripts/bench_iwae.py 2>/dev/null | tail -1 | cut -c1-900
