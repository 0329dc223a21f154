#!/bin/bash
# tensor-parallel: tests + bench at N GPUs
N=${1:-2}
O=gpurun_out
mkdir -p $O
if [ "$N" = "2" ]; then
  echo "== tp tests world 2"; timeout 700 python -m pytest tests/test_tp.py -x -q -m gpu -k world2 > $O/r2_tp2_tests.log 2>&1; echo "rc=$?"; tail -8 $O/r2_tp2_tests.log
fi
if [ "$N" = "4" ]; then
  echo "== tp tests world 4"; timeout 900 python -m pytest tests/test_tp.py -x -q -m gpu -k world4 > $O/r2_tp4_tests.log 2>&1; echo "rc=$?"; tail -8 $O/r2_tp4_tests.log
fi
echo "== bench tp$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 64 --warmup 8 > $O/r2_bench_tp$N.json 2> $O/r2_bench_tp$N.err; echo "rc=$?"; tail -3 $O/r2_bench_tp$N.err; cat $O/r2_bench_tp$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['parallelism'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
