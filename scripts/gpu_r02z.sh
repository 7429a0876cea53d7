#!/bin/bash
# round 2, call Z: threshold of the relation-free schedule at C = 192 / 384 (N*k*C elements), 4 and 8 rooms
TAG=${1:-r02z}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_dense_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "tests exit: $?"; tail -2 $O/pytest.log
for rooms in 4 8; do
for thr in 1e12 60e6 30e6 0; do
ROOMS=$rooms AOPT_RELFREE_MIN_ELEMS=$thr timeout 300 python scripts/model_step_times.py > $O/model_step_${rooms}rooms_thr$thr.txt 2>&1; echo "== $rooms rooms, min elems $thr: $(tail -3 $O/model_step_${rooms}rooms_thr$thr.txt | head -2 | tr '\n' ' ')"
done
done
