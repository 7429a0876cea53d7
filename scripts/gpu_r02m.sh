#!/bin/bash
# round 2, call M: shared-memory heap in the GRID kNN query kernel: full GPU suite, container A/B, bench A/B
TAG=${1:-r02m}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit: $?"; tail -14 $O/pytest_gpu.log
timeout 600 python scripts/knn_ab.py > $O/knn_ab.txt 2>&1; cat $O/knn_ab.txt
for v in heap list; do
AOPT_KNN_TOPK=$v timeout 600 python bench.py --steps 40 --warmup 3 --skip-e2e --no-cpu-baseline --no-gpu-reference --no-model --no-variants > $O/bench_s3dis4_$v.json 2> $O/bench_s3dis4_$v.err; python -c "
import json;d=json.load(open('$O/bench_s3dis4_$v.json'));print('$v value',d['value'],d['ms_per_step'],'knn',d['knn']);[print('  ',k['kernel'],k['ms_per_step']) for k in d['kernels'] if 'knn' in k['kernel']]"
done
