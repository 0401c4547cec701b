#!/bin/bash
# short 2-GPU re-validation after the round's later changes: parity check + one weak bench line
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/r2_multi2_check_n$N.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_multi2_bench_n$N.err > gpurun_out/r2_multi2_bench_n$N.json; tail -2 gpurun_out/r2_multi2_bench_n$N.err | grep -v "^W0\|OMP_NUM"; python scripts/show_bench.py gpurun_out/r2_multi2_bench_n$N.json | head -3
