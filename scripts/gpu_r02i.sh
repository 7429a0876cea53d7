#!/bin/bash
# round 2, call I: programmatic dependent launch inside the small-kernel chains (CSR, voxel partition, kNN grid build): A/B
TAG=${1:-r02i}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
for pdl in 1 0; do
  AOPT_PDL=$pdl timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench_pdl$pdl.txt 2>&1; echo "== pdl=$pdl"; grep -i "level\|csr\|voxel\|knn" $O/kernel_bench_pdl$pdl.txt
  AOPT_PDL=$pdl timeout 600 python bench.py --config s3dis4 --steps 150 --no-model --no-variants --no-cpu-baseline --no-gpu-reference --skip-e2e > $O/bench_s3dis4_pdl$pdl.json 2> $O/bench_s3dis4_pdl$pdl.err; head -c 260 $O/bench_s3dis4_pdl$pdl.json; echo
  AOPT_PDL=$pdl timeout 600 python bench.py --config kitti120k --steps 150 --no-model --no-variants --no-cpu-baseline --no-gpu-reference --skip-e2e > $O/bench_kitti_pdl$pdl.json 2> $O/bench_kitti_pdl$pdl.err; head -c 260 $O/bench_kitti_pdl$pdl.json; echo
done
