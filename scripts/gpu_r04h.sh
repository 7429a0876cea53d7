#!/bin/bash
# round 2, call 4H (last GPU seconds of the round): GRID kNN query kernel with a single insert site (AOPT_KNN_SITE=1, 1008 SASS
# instructions instead of 2944) — bit-exactness tests with the switch on, then the whole search timed both ways
TAG=${1:-r04h}
O=gpurun_out/$TAG
mkdir -p $O
AOPT_KNN_SITE=1 timeout 60 python -m pytest tests/test_knn_gpu.py -q -x --timeout 50 -p no:cacheprovider > $O/pytest_knn_site1.log 2>&1
echo "pytest (knn_site=1) exit: $?"; tail -2 $O/pytest_knn_site1.log
for v in 0 1; do
  AOPT_KNN_SITE=$v timeout 40 python scripts/kernel_bench.py --levels 0,1 > $O/kernel_bench_site$v.txt 2>&1
  echo "== knn_site=$v"; grep -i "level\|knn" $O/kernel_bench_site$v.txt
done
