#!/bin/bash
# round 2, call Y: skinny kernels for the patch projection / segmentation head; relation-free-everywhere A/B at 8 rooms
TAG=${1:-r02y}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_modules_gpu.py tests/test_ddp_gpu.py tests/test_callers_gpu.py tests/test_configs_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "tests exit: $?"; tail -4 $O/pytest.log
timeout 300 python scripts/model_step_times.py > $O/model_step.txt 2>&1; tail -3 $O/model_step.txt
for f in 0 1; do
ROOMS=8 AOPT_RELFREE_ALL=$f timeout 300 python scripts/model_step_times.py > $O/model_step_8rooms_relfree$f.txt 2>&1; echo "== 8 rooms AOPT_RELFREE_ALL=$f"; tail -3 $O/model_step_8rooms_relfree$f.txt
done
ROWS=70 CPU_ROWS=40 timeout 300 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; grep -n "Self C" $O/model_step_torch_profile.txt | head -2
