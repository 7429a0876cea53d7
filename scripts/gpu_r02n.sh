#!/bin/bash
# round 2, call N: bn_act / we_tail (dense.cu): parity tests, model-step A/B, torch profile of the fused step
TAG=${1:-r02n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_dense_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest_dense.log 2>&1; echo "dense exit: $?"; tail -15 $O/pytest_dense.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_pe_mlp_gpu.py tests/test_ddp_gpu.py tests/test_callers_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest_model.log 2>&1; echo "model exit: $?"; tail -15 $O/pytest_model.log
for f in 1 0; do
AOPT_FUSED_DENSE=$f timeout 300 python scripts/model_step_times.py > $O/model_step_dense$f.txt 2>&1; echo "== AOPT_FUSED_DENSE=$f"; tail -5 $O/model_step_dense$f.txt
done
ROWS=70 timeout 300 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; head -40 $O/model_step_torch_profile.txt | cut -c1-75,150-230
