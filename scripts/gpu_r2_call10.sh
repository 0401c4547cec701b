#!/bin/bash
# Round-2 call 10: IWAE step -- CUDA-graph replay, launch list (where does the GPU time go?).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== iwae eager"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_iwae_eager.err > gpurun_out/r2_iwae_eager.json; tail -3 gpurun_out/r2_iwae_eager.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_eager.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'])"
echo "== iwae graph"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline --cuda-graph 2> gpurun_out/r2_iwae_graph.err > gpurun_out/r2_iwae_graph.json; tail -5 gpurun_out/r2_iwae_graph.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_graph.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['bound_value'])"
echo "== iwae launch list (ncu durations)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_iwae_launches.csv python bench.py --workload iwae --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_iwae_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_iwae_launches.csv 2>/dev/null | head -30
echo "== hmc graph mode (1 GPU)"
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --cuda-graph 2> gpurun_out/r2_hmc_graph.err > gpurun_out/r2_hmc_graph.json; tail -2 gpurun_out/r2_hmc_graph.err; python scripts/show_bench.py gpurun_out/r2_hmc_graph.json | head -2
echo "== samplers + misc tests"
timeout 600 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_distributions.py tests/test_gpu_models.py -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -6
