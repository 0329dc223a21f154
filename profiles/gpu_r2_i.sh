#!/bin/bash
# round 2, call I: L2 prefetch depth sweep of the producer warp (tiles per warp beyond the ring)
O=gpurun_out
mkdir -p $O
TAG=${TAG:-r2i}
export CC_B200_LIB=profiles/libcc_${1:-prod}.so
CC_B200_TOKEN_PREFETCH=4 timeout 420 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${TAG}_tests.log
for pf in ${PFS:-0 2 4 8 16}; do
  CC_B200_TOKEN_PREFETCH=$pf CC_BENCH_NOCPU=1 timeout 300 python bench.py --steps 64 --warmup 8 > $O/${TAG}_pf${pf}_bench.json 2> $O/${TAG}_pf${pf}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_pf${pf}_bench.err; python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_pf${pf}_bench.json').read())
    print('PF$pf value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'])
except Exception as e: print('no bench line', e)
PY
done
for pf in ${TRACE_PFS:-4}; do
  CC_B200_TOKEN_PREFETCH=$pf timeout 200 python profiles/trace_token.py > $O/${TAG}_pf${pf}_trace.txt 2>&1; grep -v Warn $O/${TAG}_pf${pf}_trace.txt | grep -v nanmean | tail -34
done
