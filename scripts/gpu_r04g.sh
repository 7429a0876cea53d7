#!/bin/bash
# round 2, call 4G: resident CTAs per SM of gather_sub / relation_backward at the coarse levels (the key / gradient tensors
# of levels 1-3 fit L2, so the compact-window argument that fixes 3 / 4 CTAs per SM at level 0 does not apply there)
TAG=${1:-r04g}
O=gpurun_out/$TAG
mkdir -p $O
for v in "X=0" "AOPT_GATHER_SUB_CTAS=4 AOPT_RELBWD_CTAS=5" "AOPT_GATHER_SUB_CTAS=5 AOPT_RELBWD_CTAS=6" "AOPT_GATHER_SUB_CTAS=6 AOPT_RELBWD_CTAS=3" "AOPT_GATHER_SUB_CTAS=2 AOPT_RELBWD_CTAS=2"; do
  env $v timeout 300 python scripts/kernel_bench.py --levels 0,1,2,3 > "$O/kernel_bench_${v// /_}.txt" 2>&1
  echo "== kernel_bench $v"; grep -i "level\|gather_sub\|relation_backward" "$O/kernel_bench_${v// /_}.txt" | head -40
done
