#!/bin/bash
# Round-2 call 33: full GPU suite on the final defaults (dgrad from the forward weight planes on).
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_c33_pytest.log 2>&1; tail -4 gpurun_out/r2_c33_pytest.log
timeout 200 python bench.py --workload iwae --no-cpu-baseline 2> gpurun_out/r2_c33_iwae.err > gpurun_out/r2_c33_iwae.json; tail -1 gpurun_out/r2_c33_iwae.err; cut -c1-200 gpurun_out/r2_c33_iwae.json
