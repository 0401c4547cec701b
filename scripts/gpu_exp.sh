#!/bin/bash
mkdir -p gpurun_out
run() { # name impl env...
  name=$1; impl=$2; shift; shift
  env "$@" timeout 120 python bench.py --steps 3 --warmup 3 --burnin 0 --no-adapt --dense-impl $impl --no-cpu-baseline --no-e2e 2> gpurun_out/exp_$name.err > gpurun_out/exp_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-28s kernel_ms %.4f  ms/step %.2f  clocks %s' % ('$name', r['kernel_ms_per_launch'], d['ms_per_step'], d['clocks']))" 2>/dev/null || { echo "$name FAILED"; tail -3 gpurun_out/exp_$name.err; }
}
echo "== h16 parity"
timeout 300 python -m pytest tests/test_gpu_hmc.py -q -x -k "single_pass or tc_vs_simt" --no-header -p no:cacheprovider 2>&1 | tail -25
run h16 2
run tf32pair 1
