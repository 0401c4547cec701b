#!/bin/bash
# Round-2 call 2: golden replays on the TC kernels, timing experiments + ncu on the trajectory kernel.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== golden replays"
ZSB_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf -s --no-header -p no:cacheprovider -k "golden_dense64 or golden_dense1024" 2>&1 | tail -30 | tee gpurun_out/r2_golden_big.log
B="python bench.py --steps 3 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 4"
for dbg in 0 1 2 3 4 7; do
  echo "== traj dbg=$dbg"
  ZSB_TRAJ_DBG=$dbg timeout 300 $B 2> gpurun_out/r2_traj_dbg$dbg.err > gpurun_out/r2_traj_dbg$dbg.json; tail -2 gpurun_out/r2_traj_dbg$dbg.err; python scripts/show_bench.py gpurun_out/r2_traj_dbg$dbg.json | head -2
done
for cl in 8 4; do
  echo "== traj clusters=$cl"
  ZSB_TRAJ_CLUSTERS=$cl timeout 300 $B 2> gpurun_out/r2_traj_cl$cl.err > gpurun_out/r2_traj_cl$cl.json; python scripts/show_bench.py gpurun_out/r2_traj_cl$cl.json | head -2
done
echo "== ncu full (trajectory kernel, 8192 chains)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_traj -s 2 -c 1 -o gpurun_out/r2_prof_traj -f python bench.py --steps 1 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 4 --chains-per-gpu 8192 > gpurun_out/r2_ncu_traj.log 2>&1
tail -2 gpurun_out/r2_ncu_traj.log
ls -la gpurun_out/*.ncu-rep
