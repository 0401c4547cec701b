#!/bin/bash
# Multi-GPU call: sharded-vs-single parity, weak / strong scaling lines, IWAE line, reference arm.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi-gpu parity check (N=$N)"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/r2_multi_gpu_check_n$N.log
for mode in weak strong; do
  echo "== bench $mode N=$N"
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-10} --warmup 5 --scaling $mode 2> gpurun_out/r2_bench_${mode}_n$N.err > gpurun_out/r2_bench_${mode}_n$N.json; tail -3 gpurun_out/r2_bench_${mode}_n$N.err | grep -v "^W0\|OMP_NUM"; python scripts/show_bench.py gpurun_out/r2_bench_${mode}_n$N.json | head -3
done
echo "== bench strong + cuda graph N=$N"
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps ${STEPS:-10} --warmup 5 --scaling strong --cuda-graph --no-e2e 2> gpurun_out/r2_bench_strong_graph_n$N.err > gpurun_out/r2_bench_strong_graph_n$N.json; tail -3 gpurun_out/r2_bench_strong_graph_n$N.err | grep -v "^W0\|OMP_NUM"; python scripts/show_bench.py gpurun_out/r2_bench_strong_graph_n$N.json | head -2
echo "== bench iwae N=$N"
timeout 600 $TR --master-port 29515 bench.py --gpus $N --workload iwae --steps 10 --warmup 5 2> gpurun_out/r2_bench_iwae_n$N.err > gpurun_out/r2_bench_iwae_n$N.json; tail -2 gpurun_out/r2_bench_iwae_n$N.err | grep -v "^W0\|OMP_NUM"; cut -c1-330 gpurun_out/r2_bench_iwae_n$N.json
echo "== bench ref N=$N"
timeout 300 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-260
