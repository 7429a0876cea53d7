#!/bin/bash
# round 2, call D: scan look-back, kNN without the density sample (A/B), pe_mlp tc tuning, voxel pass guess -> tests, table, bench
TAG=${1:-r02d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench.txt 2>&1; grep -i "level\|pe_mlp\|csr\|voxel\|knn\|pos_mom" $O/kernel_bench.txt
AOPT_KNN_SAMPLE=sampled timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench_sampled.txt 2>&1; grep -i "level\|knn" $O/kernel_bench_sampled.txt
timeout 900 python bench.py --config s3dis4 --steps 100 > $O/bench_s3dis4.json 2> $O/bench_s3dis4.err; echo "bench s3dis4 exit: $?"; head -c 300 $O/bench_s3dis4.json; echo; tail -3 $O/bench_s3dis4.err
AOPT_KNN_SAMPLE=sampled timeout 900 python bench.py --config s3dis4 --steps 100 --no-model --no-variants --no-cpu-baseline --no-gpu-reference --skip-e2e > $O/bench_s3dis4_sampled.json 2> $O/bench_s3dis4_sampled.err; head -c 300 $O/bench_s3dis4_sampled.json; echo
for cfg in scannet150k kitti120k; do
  timeout 900 python bench.py --config $cfg --steps 100 --no-model --no-cpu-baseline --no-gpu-reference > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  echo "bench $cfg exit: $?"; head -c 300 $O/bench_$cfg.json; echo; tail -3 $O/bench_$cfg.err
done
