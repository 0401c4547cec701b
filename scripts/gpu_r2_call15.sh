#!/bin/bash
# Round-2 call 15: full GPU suite after the sampler / reinforce / meta_bn changes.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))"
timeout 1800 python -m pytest tests -x -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/r2_pytest_gpu3.log 2>&1; tail -25 gpurun_out/r2_pytest_gpu3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
