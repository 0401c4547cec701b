#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi-gpu parity check (N=$N)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/multi_gpu_check_n$N.log
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err | grep -v "^W0\|OMP_NUM"; python scripts/show_bench.py gpurun_out/bench_n$N.json
echo "== bench ref N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
