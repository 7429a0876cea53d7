#!/bin/bash
# round 2, call 4B: A/B of (a) the peb-direct fused GVA backward (AOPT_GVA_PD=1: NS smem slots per thread, 4 CTAs per SM),
# (b) relation_backward with the item's own rows bulk-prefetched into L2 (AOPT_RELBWD_PF=1); default = l2pf on for the GVA kernels
# NOTE: AOPT_GVA_PD and AOPT_RELBWD_PF selected experiment kernels that were measured slower and removed again
# (results: profiles/r04b_kernel_bench_l2pf.txt); the script is kept as the record of what was run
TAG=${1:-r04b}
O=gpurun_out/$TAG
mkdir -p $O
AOPT_GVA_PD=1 AOPT_RELBWD_PF=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_modules_gpu.py -q -x --timeout 300 -p no:cacheprovider > $O/pytest_variants.log 2>&1
echo "pytest (gva_pd=1 relbwd_pf=1) exit: $?"; tail -3 $O/pytest_variants.log
i=0
for v in "AOPT_L2PF=0" "AOPT_L2PF=1" "AOPT_GVA_PD=1" "AOPT_RELBWD_PF=1"; do
  i=$((i+1))
  env $v timeout 300 python scripts/kernel_bench.py --levels 0,1,2 > $O/kernel_bench_v$i.txt 2>&1
  echo "== kernel_bench $v"; grep -i "level\|gva_backward (\|relation_backward\|gva_forward" $O/kernel_bench_v$i.txt | head -40
done
i=0
for v in "AOPT_L2PF=1" "AOPT_GVA_PD=1" "AOPT_L2PF=1" "AOPT_GVA_PD=1"; do
  i=$((i+1))
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-model --no-cpu-baseline --no-gpu-reference --no-variants --skip-e2e > $O/bench_v$i.json 2> $O/bench.err
  echo "== bench $v exit $?"; python -c "
import sys, json
d=json.loads(open('$O/bench_v$i.json').read().strip().splitlines()[-1])
sm=sorted(d['step_ms']); print('value %.2f ms %.3f median %.3f' % (d['value'], d['ms_per_step'], sm[len(sm)//2]), [(k['kernel'], k['ms_per_step']) for k in d['kernels'][:4]])
"
done
