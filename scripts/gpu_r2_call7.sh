#!/bin/bash
# Round-2 call 7: in-place planes variant of the resident kernel; diag-kernel RNG sharing; IWAE line.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== in-place tests"
ZSB_RES_INPLACE=1 timeout 600 python -m pytest tests/test_gpu_hmc.py -m gpu -q -rf --no-header -p no:cacheprovider -k "resident_kernel or (trajectory_kernels and 5) or (golden and 5) or cuda_graph or leapfrog_count" > gpurun_out/r2_inplace_tests.log 2>&1; tail -8 gpurun_out/r2_inplace_tests.log
echo "== diag kernel tests (shared Philox blocks)"
timeout 600 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_distributions.py -m gpu -q -rf --no-header -p no:cacheprovider -k "diag or philox or golden_diag or univariate_more" 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --dense-impl 5"
for ip in 1 0 1 0; do
  echo "== sustained res INPLACE=$ip"
  ZSB_RES_INPLACE=$ip timeout 300 $B 2> gpurun_out/r2_ip$ip.err > gpurun_out/r2_ip$ip.json; tail -2 gpurun_out/r2_ip$ip.err; python scripts/show_bench.py gpurun_out/r2_ip$ip.json | head -3
done
echo "== ncu dram bytes in-place, group=37"
ZSB_RES_INPLACE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dense_res -s 2 -c 1 --csv --log-file gpurun_out/r2_l2cap_inplace.csv python bench.py --steps 1 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5 --chains-per-gpu 18944 > /dev/null 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/r2_l2cap_inplace.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
echo "== kernel microbench (diag C1')"
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_bench_kernels.jsonl 2> gpurun_out/r2_bench_kernels.err; tail -3 gpurun_out/r2_bench_kernels.err; grep -i "diag\|momentum\|normal_log_prob" gpurun_out/r2_bench_kernels.jsonl | cut -c1-300
echo "== iwae bench"
timeout 600 python bench.py --workload iwae --steps 10 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_iwae2.err > gpurun_out/r2_iwae2.json; tail -3 gpurun_out/r2_iwae2.err; cut -c1-400 gpurun_out/r2_iwae2.json
