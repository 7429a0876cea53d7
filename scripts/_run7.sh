O=gpurun_out/x7; mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "relation" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for ct in 2 3 4 6; do echo "== ctas/SM $ct"; AOPT_RELBWD_CTAS=$ct timeout 300 python scripts/kernel_bench.py --levels 0,1,2 2>&1 | grep -E "level|relation_backward"; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-cpu-baseline > $O/bench_fused.json 2>$O/bench.err
python -c "
import json
for f in ('fused',):
    d=json.load(open('$O/bench_'+f+'.json')); print(f, round(d['value'],2), round(d['ms_per_step'],3), d['hbm_kernels_total'])
    for k in d['kernels']:
        if k['kernel'] in ('aopt_relation_backward','aopt_grouping_backward','aopt_sum_over_k'): print('   ',k['kernel'],k['ms_per_step'],k['frac'])
"
