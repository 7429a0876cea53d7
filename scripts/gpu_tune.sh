#!/bin/bash
# Tuning run: kernel micro-benchmarks with variants + presort experiment.  Output in gpurun_out/$1.
TAG=${1:-tune}
O=gpurun_out/$TAG
mkdir -p $O
python scripts/kernel_bench.py --levels 0,1 > $O/kb_default.txt 2>&1
AOPT_GATHER_SUB_IMPL=rows AOPT_BV_IMPL=unroll4 python scripts/kernel_bench.py --levels 0 > $O/kb_alt.txt 2>&1
python scripts/kernel_bench.py --levels 0 --presort > $O/kb_presort.txt 2>&1
AOPT_GATHER_SUB_IMPL=rows AOPT_BV_IMPL=unroll4 python scripts/kernel_bench.py --levels 0 --presort > $O/kb_presort_alt.txt 2>&1
python bench.py --steps 6 --warmup 3 --no-model --no-cpu-baseline --skip-e2e --presort > $O/bench_presort.json 2> $O/bench_presort.err
AOPT_KNN_GRID_MIN=2048 python bench.py --steps 6 --warmup 3 --no-model --no-cpu-baseline --skip-e2e > $O/bench_gridmin2048.json 2> $O/bench_gridmin.err
timeout 600 python scripts/profile_model.py > $O/model_profile.txt 2>&1
for f in $O/kb_*.txt; do echo "== $f"; cat $f; done
python -c "
import json
for f in ('bench_presort','bench_gridmin2048'):
    d=json.load(open('$O/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['knn'])
"
head -60 $O/model_profile.txt
