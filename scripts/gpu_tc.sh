#!/bin/bash
# TC kernel bring-up: targeted tests first (short timeout: the kernel traps after ~2 s on a stuck
# barrier), then the GPU suite and bench variants.
mkdir -p gpurun_out
for BK in 32 16; do
echo "== tc single-pass + golden tests BK=$BK"
ZSB_TC_BK=$BK timeout 300 python -m pytest tests/test_gpu_hmc.py -q -x -k "single_pass or dense_fused_tc" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/tc_single_bk$BK.log
done
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
for BK in 32 16; do
echo "== bench impl1 BK=$BK"
ZSB_TC_BK=$BK timeout 600 python bench.py --steps 5 --warmup 3 --burnin 14 --dense-impl 1 --no-cpu-baseline 2> gpurun_out/bench1_bk$BK.err > gpurun_out/bench_impl1_bk$BK.json; tail -3 gpurun_out/bench1_bk$BK.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_impl1_bk$BK.json').read().strip().splitlines()[-1]); r=d['roofline']
print('BK=$BK value %.4e ms/step %.2f e2e %.4e kernel_ms %.4f share %.3f hbm_frac %.3f tensor_TF %.1f acc %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms_per_launch'], r['kernel_share_of_step'], r['frac'], r['tensor']['achieved'], d['acceptance_mean']))"
done
