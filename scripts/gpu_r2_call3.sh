#!/bin/bash
# Round-2 call 3: first run of the L2-resident trajectory kernel (impl 5) + full golden tables.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== impl 5 shapes + trajectory tests"
timeout 900 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "resident_kernel or trajectory_kernels" > gpurun_out/r2_res_tests.log 2>&1; tail -15 gpurun_out/r2_res_tests.log
echo "== golden replays (tables)"
timeout 900 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf -s --no-header -p no:cacheprovider -k "golden_dense64 or golden_dense1024" > gpurun_out/r2_golden_big.log 2>&1; grep -E "^replay|^ it|^ +[0-9]+ |passed|failed|^FAILED" gpurun_out/r2_golden_big.log | head -150
B="python bench.py --steps 3 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5"
for dbg in 0 1 2 4; do
  echo "== res dbg=$dbg"
  ZSB_RES_DBG=$dbg timeout 300 $B 2> gpurun_out/r2_res_dbg$dbg.err > gpurun_out/r2_res_dbg$dbg.json; tail -2 gpurun_out/r2_res_dbg$dbg.err; python scripts/show_bench.py gpurun_out/r2_res_dbg$dbg.json | head -3
done
for grp in 28 32 18 74; do
  echo "== res group=$grp"
  ZSB_RES_GROUP=$grp timeout 300 $B 2> gpurun_out/r2_res_grp$grp.err > gpurun_out/r2_res_grp$grp.json; tail -2 gpurun_out/r2_res_grp$grp.err; python scripts/show_bench.py gpurun_out/r2_res_grp$grp.json | head -3
done
echo "== bench impl 5 (adaptive, full)"
timeout 600 python bench.py --steps 10 --warmup 3 --dense-impl 5 --no-cpu-baseline 2> gpurun_out/r2_b_impl5.err > gpurun_out/r2_b_impl5.json; tail -3 gpurun_out/r2_b_impl5.err; python scripts/show_bench.py gpurun_out/r2_b_impl5.json
echo "== bench impl 2 (adaptive, full)"
timeout 600 python bench.py --steps 10 --warmup 3 --dense-impl 2 --no-cpu-baseline 2> gpurun_out/r2_b_impl2b.err > gpurun_out/r2_b_impl2b.json; tail -3 gpurun_out/r2_b_impl2b.err; python scripts/show_bench.py gpurun_out/r2_b_impl2b.json
echo "== ncu full (resident kernel, 18944 chains = 2 groups)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_res -s 2 -c 1 -o gpurun_out/r2_prof_res -f python bench.py --steps 1 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5 --chains-per-gpu 18944 > gpurun_out/r2_ncu_res.log 2>&1
tail -2 gpurun_out/r2_ncu_res.log
