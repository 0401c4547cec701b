#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -25
timeout 600 python scripts/bench_kernels.py 2>&1 | tail -30 | tee gpurun_out/bench_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('%-52s %8.4f ms %8.1f GB/s  frac %.3f  %s' % (d['kernel'], d['ms'], d['GBps'], d['frac_of_hbm_peak'], d['note'][:70]))
    except Exception: print(l.strip()[:200])
"
timeout 300 python scripts/bench_iwae.py 2>&1 | tail -1 | tee gpurun_out/bench_iwae.json | cut -c1-700
