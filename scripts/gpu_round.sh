#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (+ reference arm), optional ncu launch list / full capture.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -3 gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/bench_ref.json; python scripts/show_bench.py gpurun_out/bench_ref.json
if [ -n "$DO_NCU" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --burnin 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -8
echo "== ncu full (dominant kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_leapfrog_tc2 -s 20 -c 1 -o gpurun_out/prof_dom -f python bench.py --steps 1 --warmup 1 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e > gpurun_out/ncu_dom.log 2>&1
tail -1 gpurun_out/ncu_dom.log
fi
