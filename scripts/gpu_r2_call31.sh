#!/bin/bash
# Round-2 call 31: validation of the final build with the driver's command lines + microbenchmarks.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_c31_pytest.log 2>&1; tail -4 gpurun_out/r2_c31_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench.py (driver default)"
timeout 600 python bench.py 2> gpurun_out/r2_c31_bench.err > gpurun_out/r2_c31_bench.json; tail -2 gpurun_out/r2_c31_bench.err; python scripts/show_bench.py gpurun_out/r2_c31_bench.json | head -6
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2> gpurun_out/r2_c31_ref.err > gpurun_out/r2_c31_ref.json; tail -2 gpurun_out/r2_c31_ref.err; python scripts/show_bench.py gpurun_out/r2_c31_ref.json
echo "== iwae"
timeout 600 python bench.py --workload iwae 2> gpurun_out/r2_c31_iwae.err > gpurun_out/r2_c31_iwae.json; tail -2 gpurun_out/r2_c31_iwae.err; cut -c1-300 gpurun_out/r2_c31_iwae.json
echo "== kernel microbench"
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_c31_bench_kernels.jsonl 2> gpurun_out/r2_c31_bench_kernels.err; tail -3 gpurun_out/r2_c31_bench_kernels.err; grep -i "diag\|momentum\|config 1" gpurun_out/r2_c31_bench_kernels.jsonl | cut -c1-260
echo "== bnn config 4"
timeout 300 python scripts/bench_bnn.py 2>&1 | tail -4 | cut -c1-300
