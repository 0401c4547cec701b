#!/bin/bash
# Round-2 call 12: full GPU suite (incl. LNTM fused kernel, checkpoints), IWAE graph replay, LNTM timing.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== full GPU suite"
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_pytest_gpu2.log 2>&1; tail -12 gpurun_out/r2_pytest_gpu2.log
echo "== iwae graph"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline --cuda-graph 2> gpurun_out/r2_iwae_graph3.err > gpurun_out/r2_iwae_graph3.json; grep -v "^\s*$" gpurun_out/r2_iwae_graph3.err | tail -4; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_graph3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['bound_value'])"
echo "== iwae eager"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 2> gpurun_out/r2_iwae_eager3.err > gpurun_out/r2_iwae_eager3.json; python -c "
import json; d=json.loads(open('gpurun_out/r2_iwae_eager3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['cpu_baseline'])"
echo "== LNTM config-5 timing"
timeout 600 python scripts/bench_lntm.py 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
