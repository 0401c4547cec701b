#!/bin/bash
# Round-2 call 8: multicast with group sized from the real co-resident cluster count; in-place fix.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== in-place golden tests (probe fix)"
ZSB_RES_INPLACE=1 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "(golden and 5) or cuda_graph" 2>&1 | tail -4
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --dense-impl 5"
for cfg in "1 0" "1 1" "0 1" "0 0"; do
  set -- $cfg
  echo "== sustained res MC=$1 INPLACE=$2"
  ZSB_RES_VERBOSE=1 ZSB_RES_MC=$1 ZSB_RES_INPLACE=$2 timeout 300 $B 2> gpurun_out/r2_mc$1_ip$2.err > gpurun_out/r2_mc$1_ip$2.json; grep "co-resident" gpurun_out/r2_mc$1_ip$2.err | head -1; python scripts/show_bench.py gpurun_out/r2_mc$1_ip$2.json | head -3
done
echo "== MC tests again (group sizing changed)"
ZSB_RES_MC=1 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "resident_kernel or (trajectory_kernels and 5) or (golden and 5)" 2>&1 | tail -4
