#!/bin/bash
# round 2, call C: pe_mlp backward on tcgen05 (tests + table), model step in isolation (A/B of the pe kernels), ncu evidence
TAG=${1:-r02c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_pe_mlp_gpu.py -q -x --timeout 250 -p no:cacheprovider > $O/pytest_pe.log 2>&1; echo "pe_mlp exit: $?"; tail -15 $O/pytest_pe.log
timeout 300 python -m pytest tests/test_modules_gpu.py -q -x --timeout 250 -p no:cacheprovider > $O/pytest_modules.log 2>&1; echo "modules exit: $?"; tail -3 $O/pytest_modules.log
timeout 600 python scripts/kernel_bench.py --levels 0,1 > $O/kernel_bench.txt 2>&1; grep -i "level\|pe_mlp" $O/kernel_bench.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-variants --skip-e2e --no-cpu-baseline --no-gpu-reference > $O/bench_model_tc.json 2> $O/bench_model_tc.err; python -c "import json;d=json.load(open('$O/bench_model_tc.json'));print('model tc',d['model_step'])"
AOPT_PE_FWD=mma AOPT_PE_BWD=mma timeout 600 python bench.py --steps 5 --warmup 3 --no-variants --skip-e2e --no-cpu-baseline --no-gpu-reference > $O/bench_model_mma.json 2> $O/bench_model_mma.err; python -c "import json;d=json.load(open('$O/bench_model_mma.json'));print('model mma',d['model_step'])"
if [ "${SKIP_NCU:-0}" != "1" ]; then bash scripts/gpu_ncu.sh $TAG; fi
