#!/bin/bash
# round 2, call E (N GPUs of one box): scene-sharded schedule + DDP-wrapped model step under torchrun, NCCL log
N=${N:-2}
TAG=${1:-r02e}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $O/gpu_info.csv 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -x --timeout 500 -p no:cacheprovider > $O/pytest_ddp.log 2>&1; echo "ddp test exit: $?"; tail -3 $O/pytest_ddp.log
run() {  # config steps extra...
  cfg=$1; st=$2; shift 2
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=$O/nccl_${cfg}_n$N.%h.%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps $st --warmup 3 "$@" > $O/bench_${cfg}_n$N.json 2> $O/bench_${cfg}_n$N.err
  echo "bench $cfg n=$N exit: $?"; head -c 300 $O/bench_${cfg}_n$N.json; echo; tail -2 $O/bench_${cfg}_n$N.err
}
run s3dis4 100 --no-variants
run s3dis8 60 --no-variants
# one-GPU lines of the same box for the scaling ratio
timeout 600 python bench.py --config s3dis4 --steps 100 --no-variants --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis4_n1.json 2> $O/bench_s3dis4_n1.err; head -c 200 $O/bench_s3dis4_n1.json; echo
timeout 600 python bench.py --config s3dis8 --steps 60 --no-variants --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis8_n1.json 2> $O/bench_s3dis8_n1.err; head -c 200 $O/bench_s3dis8_n1.json; echo
timeout 600 python scripts/profile_model.py > $O/model_step_torch_profile.txt 2>&1; head -30 $O/model_step_torch_profile.txt | cut -c1-180
# NCCL: algorithm / protocol / channels actually used by the gradient all-reduce
cat $O/nccl_*.log 2>/dev/null | grep -i "AllReduce\|NVLS\|Channel\|algo\|proto\|nvlink\|P2P" | sort | uniq -c | sort -rn | head -40 > $O/nccl_summary.txt
rm -f $O/nccl_*.log
head -20 $O/nccl_summary.txt
