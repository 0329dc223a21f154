#!/bin/bash
# 2-GPU check: TP tests, bench with the fused NVLink exchange and with NCCL all-reduces.  gpurun --gpus 2
TAG=${1:-r1}
N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_tp_gpus.txt
echo "== pytest tp"; timeout 600 python -m pytest tests/test_tp.py -x -q -m gpu > $O/${TAG}_pytest_tp.log 2>&1; echo "rc=$?"; tail -5 $O/${TAG}_pytest_tp.log
for mode in fused nccl; do
  echo "== bench --gpus $N ($mode)"
  if [ $mode = nccl ]; then export CC_TP_NCCL=1; else unset CC_TP_NCCL; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 64 --warmup 8 > $O/${TAG}_bench_tp${N}_${mode}.json 2> $O/${TAG}_bench_tp${N}_${mode}.err
  echo "rc=$?"; cat $O/${TAG}_bench_tp${N}_${mode}.json; tail -3 $O/${TAG}_bench_tp${N}_${mode}.err
done
