#!/bin/bash
# round 2, call 4F: state at the end of the round — full GPU suite, smoke, default bench of the four configs, reference arm,
# per-operator table (ncu evidence: scripts/gpu_ncu.sh in its own call)
TAG=${1:-r04f}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu_info.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit: $?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_s3dis4.json 2> $O/bench_s3dis4.err; echo "bench s3dis4 exit: $?"; head -c 300 $O/bench_s3dis4.json; echo; tail -3 $O/bench_s3dis4.err
for cfg in s3dis8 scannet150k kitti120k; do
  timeout 600 python bench.py --config $cfg --steps 100 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  echo "bench $cfg exit: $?"; head -c 300 $O/bench_$cfg.json; echo; tail -3 $O/bench_$cfg.err
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 300 $O/bench_ref.json; echo
timeout 300 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench.txt 2>&1; grep -i "level\|gva\|relation\|gather_sub\|knn\|csr" $O/kernel_bench.txt
