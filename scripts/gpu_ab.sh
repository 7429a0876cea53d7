#!/bin/bash
# One gpurun call: GPU parity tests, then A/B bench runs of kernel variants selected by AOPT_* variables.
# Output in gpurun_out/$1.  usage: scripts/gpu_ab.sh TAG ["VAR=val VAR2=val" ...]   (each quoted arg = one variant)
TAG=${1:-ab}; shift
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_info.csv 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider --durations=12 ${PYTEST_ARGS:-} > $O/pytest_gpu.log 2>&1
  echo "pytest exit: $?" >> $O/pytest_gpu.log
  tail -25 $O/pytest_gpu.log
fi
summ() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.2f Mpts/s  %.3f ms/step  e2e %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for k in d["kernels"]:
    print("   %-30s %7.3f ms  x%-3d frac %-6s largest %7.1f us frac %s" % (k["kernel"], k["ms_per_step"], k["calls_per_step"], k["frac"], k["largest"]["us_per_launch"], k["largest"]["frac"]))
PY
}
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-model --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"; summ $O/bench_default.json
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-model --no-cpu-baseline --skip-e2e > $O/bench_v$i.json 2> $O/bench_v$i.err
  echo "== variant $i: $v"; summ $O/bench_v$i.json
done
if [ "${KB:-1}" = "1" ]; then
  timeout 600 python scripts/kernel_bench.py --levels 0 > $O/kernel_bench_L0.txt 2>&1; cat $O/kernel_bench_L0.txt
fi
