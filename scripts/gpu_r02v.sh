#!/bin/bash
# round 2, call V: skinny Linear kernels + relation-free schedule at every width: parity tests, model-step A/B, profile
TAG=${1:-r02v}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_dense_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest_dense.log 2>&1; echo "dense exit: $?"; tail -5 $O/pytest_dense.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_ddp_gpu.py tests/test_callers_gpu.py tests/test_pe_mlp_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest_model.log 2>&1; echo "model exit: $?"; tail -5 $O/pytest_model.log
for f in 1 0; do
AOPT_RELFREE_ALL=$f timeout 300 python scripts/model_step_times.py > $O/model_step_relfree$f.txt 2>&1; echo "== AOPT_RELFREE_ALL=$f"; tail -4 $O/model_step_relfree$f.txt
done
ROWS=60 CPU_ROWS=45 timeout 300 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; grep -n "Self C" $O/model_step_torch_profile.txt | head -2
