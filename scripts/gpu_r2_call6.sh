#!/bin/bash
# Round-2 call 6: where does the L2 stop holding a chain group?  DRAM bytes of the resident kernel vs group size.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
for grp in 12 18 24 30 37; do
  echo "== ncu dram bytes, group=$grp"
  ZSB_RES_GROUP=$grp timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dense_res -s 2 -c 1 --csv --log-file gpurun_out/r2_l2cap_$grp.csv python bench.py --steps 1 --warmup 3 --burnin 0 --no-adapt --no-cpu-baseline --no-e2e --dense-impl 5 --chains-per-gpu 18944 > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r2_l2cap_$grp.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
echo "== AIS device loop + samplers tests"
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_samplers.py tests/test_gpu_distributions.py -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -15
