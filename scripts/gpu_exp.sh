#!/bin/bash
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 3 --warmup 3 --burnin 0 --no-adapt --dense-impl 1 --no-cpu-baseline --no-e2e 2> gpurun_out/exp_$name.err > gpurun_out/exp_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-28s kernel_ms %.4f  ms/step %.2f' % ('$name', r['kernel_ms_per_launch'], d['ms_per_step']))" 2>/dev/null || { echo "$name FAILED"; tail -3 gpurun_out/exp_$name.err; }
}
for BK in 32 16; do
echo "== pair kernel parity BK=$BK"
ZSB_TC_PAIR=1 ZSB_TC_BK=$BK timeout 200 python -m pytest tests/test_gpu_hmc.py -q -x -k "single_pass or dense_fused_tc or tc_vs_simt" --no-header -p no:cacheprovider 2>&1 | tail -12
done
run pair32 ZSB_TC_PAIR=1 ZSB_TC_BK=32
run pair16 ZSB_TC_PAIR=1 ZSB_TC_BK=16
run single32 ZSB_TC_BK=32
