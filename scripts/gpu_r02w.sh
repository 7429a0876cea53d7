#!/bin/bash
# round 2, call W: row-unrolled BatchNorm kernels + linear_bn_act: parity tests, per-operator table, model-step timing, profile
TAG=${1:-r02w}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_modules_gpu.py tests/test_ddp_gpu.py tests/test_callers_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "tests exit: $?"; tail -4 $O/pytest.log
timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench.txt 2>&1; grep -i "level\|bn_act\|we_tail" $O/kernel_bench.txt
timeout 300 python scripts/model_step_times.py > $O/model_step.txt 2>&1; tail -4 $O/model_step.txt
ROWS=70 CPU_ROWS=40 timeout 300 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; grep -n "Self C" $O/model_step_torch_profile.txt | head -2
