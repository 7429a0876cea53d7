#!/bin/bash
# round 2, call X (2 GPUs of one box): scene-sharded schedule + DDP-wrapped model step (fused dense path) under torchrun
N=${N:-2}
TAG=${1:-r02x}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -x --timeout 500 -p no:cacheprovider > $O/pytest_ddp.log 2>&1; echo "ddp test exit: $?"; tail -3 $O/pytest_ddp.log
run() {
  cfg=$1; st=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --config $cfg --steps $st --warmup 3 --no-variants --no-cpu-baseline --no-gpu-reference > $O/bench_${cfg}_n$N.json 2> $O/bench_${cfg}_n$N.err
  echo "bench $cfg n=$N exit: $?"; head -c 250 $O/bench_${cfg}_n$N.json; echo; tail -2 $O/bench_${cfg}_n$N.err
}
run s3dis4 60
run s3dis8 40
timeout 600 python bench.py --config s3dis4 --steps 60 --no-variants --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis4_n1.json 2> $O/bench_s3dis4_n1.err; head -c 200 $O/bench_s3dis4_n1.json; echo
timeout 600 python bench.py --config s3dis8 --steps 40 --no-variants --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis8_n1.json 2> $O/bench_s3dis8_n1.err; head -c 200 $O/bench_s3dis8_n1.json; echo
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02x/bench_*.json')):
    d=json.load(open(f)); ms=d.get('model_step') or {}
    print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f model_step %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], ms.get('ms_per_step', -1)))
PY
