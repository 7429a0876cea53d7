#!/bin/bash
# round 2, call 3B: split-K (bf16-result) weight gradients for the huge-K products: tests, model step at 4 / 8 rooms, profile
TAG=${1:-r03h}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_modules_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "tests exit: $?"; tail -2 $O/pytest.log
timeout 300 python scripts/model_step_times.py > $O/model_step.txt 2>&1; tail -3 $O/model_step.txt
ROOMS=8 timeout 300 python scripts/model_step_times.py > $O/model_step_8rooms.txt 2>&1; tail -3 $O/model_step_8rooms.txt
ROWS=70 CPU_ROWS=40 timeout 300 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; grep -n "Self C" $O/model_step_torch_profile.txt | head -2
