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
run pair32 ZSB_TC_PAIR=1 ZSB_TC_BK=32
run pair32_noepi ZSB_TC_PAIR=1 ZSB_TC_BK=32 ZSB_TC_DBG=1
run pair32_onemma ZSB_TC_PAIR=1 ZSB_TC_BK=32 ZSB_TC_DBG=2
run pair32_noepi_onemma ZSB_TC_PAIR=1 ZSB_TC_BK=32 ZSB_TC_DBG=3
run pair16_noepi ZSB_TC_PAIR=1 ZSB_TC_BK=16 ZSB_TC_DBG=1
echo "== ncu pair32"
ZSB_TC_PAIR=1 ZSB_TC_BK=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_leapfrog_tc2 -s 20 -c 1 -o gpurun_out/prof_tc2 -f python bench.py --steps 1 --warmup 1 --burnin 0 --no-adapt --dense-impl 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc2.log 2>&1
tail -2 gpurun_out/ncu_tc2.log
