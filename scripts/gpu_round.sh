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
  ls -la $O
fi
