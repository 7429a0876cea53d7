#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture of the
# hot-path kernels.  Everything lands in gpurun_out/ (tag = $1, default "r1").
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu_info.csv 2>&1
nproc > $O/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
  echo "pytest exit: $?" >> $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit: $?" >> $O/smoke.log; tail -2 $O/smoke.log
fi
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit: $?"; tail -c 3000 $O/bench.json; tail -5 $O/bench.err
if [ "${SKIP_REF:-0}" != "1" ]; then
  AOPT_BENCH_CPU_BUDGET=60 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of one bench step (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-model --no-cpu-baseline --skip-e2e --min-warmup 1 > $O/bench_under_ncu.json 2> $O/ncu_launch.err
  # full capture of every hot-path kernel at level-0 shapes
  timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -c 80 \
      -o $O/ops_L0 -f python scripts/profile_ops.py 0 > $O/ncu_full.log 2>&1
  tail -2 $O/ncu_full.log
  # gpurun_out/ is capped at 64 MiB: export what we read (raw metrics + per-kernel source pages) and drop the report
  ncu -i $O/ops_L0.ncu-rep --page raw --csv > $O/ops_L0_raw.csv 2> /dev/null
  for kname in gva_forward_ns_kernel gva_backward_query_ns_kernel csr_walk_kernel gather_sub_ns_kernel relation_backward_vec_kernel knn_grid_kernel group_xyz_packed_kernel csr_rank_kernel pool_forward_kernel interp_forward_kernel; do
    ncu -i $O/ops_L0.ncu-rep --page source --csv -k regex:$kname -c 1 > $O/src_$kname.csv 2> /dev/null
  done
  rm -f $O/ops_L0.ncu-rep
  gzip -f $O/src_*.csv $O/launches.csv
  du -sh $O
fi
if [ "${KNN_SWEEP:-0}" = "1" ]; then
  timeout 600 python scripts/knn_sweep.py > $O/knn_sweep.txt 2>&1; cat $O/knn_sweep.txt
fi
if [ "${MAKE_GOLDEN:-0}" = "1" ]; then
  python tests/golden/make_knn_golden_gpu.py $O/knn_ref_cuda.npz
fi
