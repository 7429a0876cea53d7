#!/bin/bash
# round 2, call A: new kernels first (short timeouts), GPU parity tests, kernel table, the four bench configs, reference arm
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu_info.csv 2>&1
nproc > $O/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  # the kernels written this round, each under its own short limit (a hang must not eat the call)
  timeout 240 python -m pytest tests/test_pe_mlp_gpu.py -q -x --timeout 200 -p no:cacheprovider > $O/pytest_pe.log 2>&1; echo "pe_mlp(tcgen05) exit: $?"; tail -4 $O/pytest_pe.log
  timeout 400 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -p no:cacheprovider -k "csr or voxel or grid_pool or fused" > $O/pytest_new.log 2>&1; echo "new kernels exit: $?"; tail -6 $O/pytest_new.log
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
  echo "pytest exit: $?" >> $O/pytest_gpu.log
  tail -12 $O/pytest_gpu.log
fi
timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench.txt 2>&1; grep -i "gva_backward\|level\|pe_mlp\|csr\|voxel" $O/kernel_bench.txt
AOPT_CSR_IMPL=count AOPT_PE_FWD=mma timeout 600 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench_old.txt 2>&1; grep -i "level\|pe_mlp\|csr" $O/kernel_bench_old.txt
for cfg in ${CONFIGS:-s3dis4 scannet150k kitti120k s3dis8}; do
  st=${STEPS:-60}
  timeout 900 python bench.py --config $cfg --steps $st --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  echo "bench $cfg exit: $?"; head -c 700 $O/bench_$cfg.json; echo; tail -3 $O/bench_$cfg.err
done
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 400 $O/bench_ref.json
fi
