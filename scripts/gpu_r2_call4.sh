#!/bin/bash
# Round-2 call 4: full GPU suite, sustained-mode group-size sweep of the resident kernel, IWAE bench line.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
nvidia-smi -q -d POWER | grep -E "Power Limit|Power Draw" | head -8
echo "== full GPU suite"
timeout 1800 python -m pytest tests -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; grep -E "^replay|^ it|^ +[0-9]+ +[0-9]\.[0-9]+ |passed|failed|^FAILED" gpurun_out/r2_pytest_gpu.log | head -150
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --dense-impl 5"
for grp in 37 74 148 256; do
  echo "== sustained res group=$grp"
  ZSB_RES_GROUP=$grp timeout 300 $B 2> gpurun_out/r2_sus_grp$grp.err > gpurun_out/r2_sus_grp$grp.json; tail -2 gpurun_out/r2_sus_grp$grp.err; python scripts/show_bench.py gpurun_out/r2_sus_grp$grp.json | head -3
done
echo "== sustained res dbg=1 (no epilogue memory)"
ZSB_RES_DBG=1 timeout 300 python bench.py --steps 10 --warmup 5 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5 2> gpurun_out/r2_sus_dbg1.err > gpurun_out/r2_sus_dbg1.json; python scripts/show_bench.py gpurun_out/r2_sus_dbg1.json | head -3
echo "== sustained impl 2"
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --dense-impl 2 2> gpurun_out/r2_sus_impl2.err > gpurun_out/r2_sus_impl2.json; python scripts/show_bench.py gpurun_out/r2_sus_impl2.json | head -3
echo "== iwae bench"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 3 2> gpurun_out/r2_iwae.err > gpurun_out/r2_iwae.json; tail -3 gpurun_out/r2_iwae.err; cut -c1-1500 gpurun_out/r2_iwae.json
