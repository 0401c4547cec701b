#!/bin/bash
# experiment: where does the dense kernel's time go under the power cap? (dbg: 1 = no epilogue,
# 2 = one MMA product instead of three, 3 = both)
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
echo "== dbg $dbg"
ZSB_TC_DBG=$dbg timeout 200 python bench.py --steps 10 --warmup 3 --no-adapt --burnin 2 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/b_dbg$dbg.json; python scripts/show_bench.py gpurun_out/b_dbg$dbg.json | head -3
done
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err > gpurun_out/bench.json; python scripts/show_bench.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
