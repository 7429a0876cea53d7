#!/bin/bash
# round 2, call B: changed kernels (pe_mlp tcgen05 epilogue, NS=32 blocks, kMaxRing, CSR default), kernel table, benches
TAG=${1:-r02b}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu_info.csv 2>&1
timeout 240 python -m pytest tests/test_pe_mlp_gpu.py -q -x --timeout 200 -p no:cacheprovider > $O/pytest_pe.log 2>&1; echo "pe_mlp exit: $?"; tail -3 $O/pytest_pe.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log; grep "bf16 vs fp32" $O/pytest_gpu.log
timeout 600 python -m pytest tests/test_modules_gpu.py -q -s -k bf16 -p no:cacheprovider 2>&1 | grep "bf16 vs" > $O/bf16_tolerance.txt; cat $O/bf16_tolerance.txt
timeout 600 python scripts/kernel_bench.py --levels 0,1 > $O/kernel_bench.txt 2>&1; grep -i "gva_backward\|level\|pe_mlp\|csr\|voxel\|knn" $O/kernel_bench.txt
timeout 900 python bench.py --config s3dis4 > $O/bench_s3dis4.json 2> $O/bench_s3dis4.err; echo "bench s3dis4 exit: $?"; head -c 400 $O/bench_s3dis4.json; echo; tail -3 $O/bench_s3dis4.err
for cfg in scannet150k kitti120k; do
  timeout 900 python bench.py --config $cfg --steps 100 --no-model > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  echo "bench $cfg exit: $?"; head -c 300 $O/bench_$cfg.json; echo; tail -3 $O/bench_$cfg.err
done
