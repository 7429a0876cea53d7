#!/bin/bash
# round 2, call F (N GPUs): scene-sharded schedule + DDP-wrapped model step under torchrun (short limits), NCCL log
N=${N:-2}
TAG=${1:-r02f}
O=gpurun_out/$TAG
mkdir -p $O
run() {  # config steps extra...
  cfg=$1; st=$2; shift 2
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=$O/nccl_${cfg}_n$N.%h.%p.log timeout ${LIMIT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps $st --warmup 3 "$@" > $O/bench_${cfg}_n$N.json 2> $O/bench_${cfg}_n$N.err
  echo "bench $cfg n=$N exit: $?"; head -c 300 $O/bench_${cfg}_n$N.json; echo; tail -2 $O/bench_${cfg}_n$N.err | cut -c1-300
}
run s3dis4 100 --no-variants
run s3dis8 60 --no-variants
if [ "${WITH_N1:-0}" = "1" ]; then   # one-GPU lines of the same box (scaling ratio)
  timeout 300 python bench.py --config s3dis4 --steps 100 --warmup 3 --no-variants --no-model --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis4_n1.json 2> $O/bench_s3dis4_n1.err; head -c 200 $O/bench_s3dis4_n1.json; echo
  timeout 300 python bench.py --config s3dis8 --steps 60 --warmup 3 --no-variants --no-model --no-cpu-baseline --no-gpu-reference > $O/bench_s3dis8_n1.json 2> $O/bench_s3dis8_n1.err; head -c 200 $O/bench_s3dis8_n1.json; echo
fi
cat $O/nccl_*.log 2>/dev/null | grep -i "AllReduce: [0-9]\|NVLS\|via P2P\|Algo\|nranks" | sed 's/^[^ ]* //' | sed 's/0x[0-9a-f]*/PTR/g' | sort | uniq -c | sort -rn | head -40 > $O/nccl_summary.txt
rm -f $O/nccl_*.log
head -12 $O/nccl_summary.txt
