#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py 2>&1 | tail -30 | tee gpurun_out/bench_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('%-52s %8.4f ms %8.1f GB/s  frac %.3f  %s' % (d['kernel'], d['ms'], d['GBps'], d['frac_of_hbm_peak'], d['note'][:70]))
    except Exception: print(l.strip()[:200])
"
echo "== bnn + iwae quick"
timeout 300 python scripts/bench_bnn.py 2>&1 | tail -2 | cut -c1-160
timeout 300 python scripts/bench_iwae.py 2>&1 | tail -1 | cut -c1-400
