#!/bin/bash
# round 2, call 4A: bulk L2 prefetch of the next work item (tuning "l2pf") in gva_forward / gva_backward (fused) /
# relation_backward — parity with the switch on, then A/B per operator and per step; tester vote kernel test
# NOTE: at the time of this call AOPT_L2PF=1 also switched the next-item prefetch on in relation_backward; that part was
# measured slower and removed — today the switch covers the GVA kernels only (default on, AOPT_L2PF=0 = off)
TAG=${1:-r04a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_info.csv 2>&1
AOPT_L2PF=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_callers_gpu.py tests/test_modules_gpu.py -q -x --timeout 300 -p no:cacheprovider > $O/pytest_l2pf1.log 2>&1
echo "pytest (l2pf=1) exit: $?"; tail -3 $O/pytest_l2pf1.log
for v in 0 1; do
  AOPT_L2PF=$v timeout 300 python scripts/kernel_bench.py --levels 0,1 > $O/kernel_bench_l2pf$v.txt 2>&1
  echo "== kernel_bench l2pf=$v"; grep -i "level\|gva\|relation" $O/kernel_bench_l2pf$v.txt | head -40
done
for v in 0 1 0 1; do
  AOPT_L2PF=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-model --no-cpu-baseline --no-gpu-reference --no-variants --skip-e2e > $O/bench_l2pf${v}_$RANDOM.json 2> $O/bench.err
  echo "== bench l2pf=$v exit $?"; cat $(ls -t $O/bench_l2pf${v}_*.json | head -1) | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line); print('value %.2f ms %.3f' % (d['value'], d['ms_per_step']), [(k['kernel'], k['ms_per_step']) for k in d['kernels'][:5]])
" | tail -1
done
