#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, kernel launch list.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import tensorflow" > gpurun_out/tf_probe.txt 2>&1; echo "tf import rc=$?" >> gpurun_out/tf_probe.txt
nproc >> gpurun_out/gpu.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:---steps 3 --warmup 3 --burnin 12} 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
if [ -n "$DO_NCU" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --burnin 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
fi
