#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (+ reference arm), optional ncu launch list.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== pytest (1-CTA TC variant, BK=16)"; ZSB_TC_PAIR=0 ZSB_TC_BK=16 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -k "dense" --no-header -p no:cacheprovider 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -3 gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/bench_ref.json; python scripts/show_bench.py gpurun_out/bench_ref.json
if [ -n "$DO_NCU" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --burnin 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | head -8
fi
