#!/bin/bash
# round-end dry run on one GPU: full GPU suite, smoke, the driver's two bench command lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_call23_tests.log
cat gpurun_out/r2_call23_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_call23_ref.json 2> gpurun_out/r2_call23_ref.err; cut -c1-300 gpurun_out/r2_call23_ref.json; tail -2 gpurun_out/r2_call23_ref.err
timeout 900 python bench.py > gpurun_out/r2_call23_bench.json 2> gpurun_out/r2_call23_bench.err; cat gpurun_out/r2_call23_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline'])"; tail -2 gpurun_out/r2_call23_bench.err
