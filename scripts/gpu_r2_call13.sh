#!/bin/bash
# Round-2 call 13 (8 GPUs): the north_star's strong-scaling point, weak scaling, IWAE at 8.
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in strong weak; do
  echo "== bench $mode N=$N"
  timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 5 --scaling $mode --no-cpu-baseline 2> gpurun_out/r2_bench_${mode}_n$N.err > gpurun_out/r2_bench_${mode}_n$N.json; tail -3 gpurun_out/r2_bench_${mode}_n$N.err | grep -v "^W0\|OMP_NUM\|^\*\*"; python scripts/show_bench.py gpurun_out/r2_bench_${mode}_n$N.json | head -3
done
echo "== bench iwae N=$N"
timeout 400 $TR --master-port 29515 bench.py --gpus $N --workload iwae --steps 10 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench_iwae_n$N.err > gpurun_out/r2_bench_iwae_n$N.json; tail -2 gpurun_out/r2_bench_iwae_n$N.err | grep -v "^W0\|OMP_NUM\|^\*\*"; cut -c1-330 gpurun_out/r2_bench_iwae_n$N.json
