#!/bin/bash
# N-GPU bench of the fused exchange (and optionally the TP tests / the NCCL arm).  gpurun --gpus N -- bash profiles/gpu_tpN.sh tag N [tests] [nccl]
TAG=$1; N=$2; O=gpurun_out; mkdir -p $O
if [[ " $* " == *" tests "* ]]; then
  timeout 600 python -m pytest tests/test_tp.py -x -q -m gpu > $O/${TAG}_pytest_tp.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_tp.log
fi
modes="fused"; if [[ " $* " == *" nccl "* ]]; then modes="fused nccl"; fi
for mode in $modes; do
  if [ $mode = nccl ]; then export CC_TP_NCCL=1; else unset CC_TP_NCCL; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 64 --warmup 8 > $O/${TAG}_bench_tp${N}_${mode}.json 2> $O/${TAG}_bench_tp${N}_${mode}.err
  echo "$mode rc=$?"; python -c "
import json,sys
d=json.loads(open('$O/${TAG}_bench_tp${N}_${mode}.json').read()); print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parallelism'])"
done
