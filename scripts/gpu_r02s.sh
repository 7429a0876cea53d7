#!/bin/bash
# round 2, call S: pending-list kNN query: kNN / config parity tests, variant A/B, bench A/B
TAG=${1:-r02s}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_configs_gpu.py tests/test_fullsize_gpu.py tests/test_dense_gpu.py -q -x --timeout 600 -p no:cacheprovider > $O/pytest_knn.log 2>&1; echo "knn exit: $?"; tail -4 $O/pytest_knn.log
timeout 600 python scripts/knn_ab.py > $O/knn_ab.txt 2>&1; cat $O/knn_ab.txt
for v in 1 0; do
AOPT_KNN_PEND=$v timeout 600 python bench.py --steps 40 --warmup 3 --skip-e2e --no-cpu-baseline --no-gpu-reference --no-model --no-variants > $O/bench_s3dis4_pend$v.json 2> $O/bench_s3dis4_pend$v.err; python -c "
import json;d=json.load(open('$O/bench_s3dis4_pend$v.json'));print('pend=$v value',d['value'],d['ms_per_step']);[print('  ',k['kernel'],k['ms_per_step']) for k in d['kernels'] if 'knn' in k['kernel']]"
done
