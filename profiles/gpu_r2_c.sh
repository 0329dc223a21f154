#!/bin/bash
O=gpurun_out
mkdir -p $O
TAG=${1:-r2e}
echo "== token kernel tests"; timeout 900 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu > $O/${TAG}_tok.log 2>&1; echo "rc=$?"; tail -5 $O/${TAG}_tok.log
echo "== trace"; timeout 300 python profiles/trace_token.py > $O/${TAG}_trace.txt 2>&1; tail -12 $O/${TAG}_trace.txt
echo "== bench"; CC_BENCH_NOCPU=1 timeout 600 python bench.py --steps 64 --warmup 8 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "rc=$?"; python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench.json').read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'])
PY
