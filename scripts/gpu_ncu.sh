#!/bin/bash
# ncu evidence of one round: launch list of a bench step (shares) + `--set full` capture of every hot-path kernel at
# level-0 shapes (scripts/profile_ops.py).  Numbers printed by runs under ncu are never bench values.
TAG=${1:-r02n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-model --no-cpu-baseline --no-gpu-reference --no-variants --skip-e2e --min-warmup 1 > $O/bench_under_ncu.json 2> $O/ncu_launch.err
python scripts/launch_shares.py $O/launches.csv > $O/launch_shares.md 2>/dev/null; head -30 $O/launch_shares.md
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -c 120 \
    -o $O/ops_L0 -f python scripts/profile_ops.py 0 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ncu -i $O/ops_L0.ncu-rep --page raw --csv > $O/ops_L0_raw.csv 2> /dev/null
for kname in gva_backward_fused_ns_kernel gva_forward_ns_kernel pe_mlp_forward_tc_kernel pe_mlp_backward_tc_kernel knn_grid_kernel bn_bwd_apply_kernel we_bwd_partial_kernel; do
  ncu -i $O/ops_L0.ncu-rep --page source --csv -k regex:$kname -c 1 > $O/src_$kname.csv 2> /dev/null
done
rm -f $O/ops_L0.ncu-rep
gzip -f $O/src_*.csv $O/launches.csv
python scripts/ncu_summary.py $O/ops_L0_raw.csv > $O/ops_L0_ncu_summary.md 2>/dev/null; head -50 $O/ops_L0_ncu_summary.md
du -sh $O
