#!/bin/bash
# Timing experiments + ncu capture of the TC kernel.
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --burnin 2 --dense-impl 1 --no-cpu-baseline --no-e2e 2> gpurun_out/exp_$name.err > gpurun_out/exp_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-28s kernel_ms %.4f  ms/step %.2f' % ('$name', r['kernel_ms_per_launch'], d['ms_per_step']))" || tail -3 gpurun_out/exp_$name.err
}
run base32 ZSB_TC_BK=32
run noepi32 ZSB_TC_BK=32 ZSB_TC_DBG=1
run onemma32 ZSB_TC_BK=32 ZSB_TC_DBG=2
run noepi_onemma32 ZSB_TC_BK=32 ZSB_TC_DBG=3
run noepi_onemma_nolo32 ZSB_TC_BK=32 ZSB_TC_DBG=7
run noepi16 ZSB_TC_BK=16 ZSB_TC_DBG=1
run noepi_onemma_nolo16 ZSB_TC_BK=16 ZSB_TC_DBG=7
echo "== ncu full"
ZSB_TC_BK=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_leapfrog_tc -s 20 -c 2 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 1 --burnin 1 --dense-impl 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log; ls -la gpurun_out/*.ncu-rep
