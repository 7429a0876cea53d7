#!/bin/bash
# round 2, call K: pe_mlp tcgen05 kernels with the one-tile-ahead prefetch: tests + table
TAG=${1:-r02k}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_pe_mlp_gpu.py tests/test_modules_gpu.py -q -x --timeout 250 -p no:cacheprovider > $O/pytest_pe.log 2>&1; echo "pe_mlp + modules exit: $?"; tail -4 $O/pytest_pe.log
timeout 600 python scripts/kernel_bench.py --levels 0,1 > $O/kernel_bench.txt 2>&1; grep -i "level\|pe_mlp\|pos_mom" $O/kernel_bench.txt
timeout 600 python bench.py --steps 30 --warmup 3 --skip-e2e --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis4.json 2> $O/bench_s3dis4.err; python -c "
import json;d=json.load(open('$O/bench_s3dis4.json'));print('value',d['value'],'fused',d['variants']['fused']['value'],d['variants']['fused']['ms_per_step'],'model',d['model_step']['ms_per_step'])"
