#!/bin/bash
# Round-2 call 14: evidence for the final build -- driver command, reference arm, launch list, ncu full.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
echo "== driver command: bench.py --gpus 1 --steps 20 --warmup 5"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_final_n1.err > gpurun_out/r2_final_n1.json; tail -2 gpurun_out/r2_final_n1.err; python scripts/show_bench.py gpurun_out/r2_final_n1.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 2>/dev/null > gpurun_out/r2_final_ref.json; python scripts/show_bench.py gpurun_out/r2_final_ref.json
echo "== iwae default + reference arm"
timeout 600 python bench.py --workload iwae --steps 20 --warmup 5 2> gpurun_out/r2_final_iwae.err > gpurun_out/r2_final_iwae.json; cut -c1-200 gpurun_out/r2_final_iwae.json
timeout 600 python bench.py --workload iwae --impl reference --steps 3 2>/dev/null > gpurun_out/r2_final_iwae_ref.json; cut -c1-160 gpurun_out/r2_final_iwae_ref.json
echo "== ncu launch list of one HMC benchmark step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_hmc.csv python bench.py --steps 2 --warmup 1 --burnin 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_hmc.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_hmc.csv 2>/dev/null | head -14
echo "== ncu full (dense_res_kernel, 65536 chains)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_res -s 2 -c 1 -o gpurun_out/r2_prof_res_final -f python bench.py --steps 1 --warmup 2 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_res_final.log 2>&1
tail -1 gpurun_out/r2_ncu_res_final.log | cut -c1-200
ncu -i gpurun_out/r2_prof_res_final.ncu-rep --page raw --csv > gpurun_out/r2_prof_res_final_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_prof_res_final.ncu-rep | head -30
echo "== LNTM"
timeout 600 python scripts/bench_lntm.py 2>/dev/null | tail -1 > gpurun_out/r2_lntm.json; cat gpurun_out/r2_lntm.json | cut -c1-400
