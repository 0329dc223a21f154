#!/bin/bash
O=gpurun_out
mkdir -p $O
TAG=${1:-r2f}
echo "== token kernel tests"; timeout 900 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu > $O/${TAG}_tok.log 2>&1; echo "rc=$?"; tail -3 $O/${TAG}_tok.log
for pf in 0 2 4 8; do
  echo "== prefetch $pf"
  CC_B200_TOKEN_PREFETCH=$pf timeout 300 python profiles/trace_token.py > $O/${TAG}_trace_pf$pf.txt 2>&1; tail -10 $O/${TAG}_trace_pf$pf.txt | grep -v Warn | grep -v nanmean
  CC_B200_TOKEN_PREFETCH=$pf CC_BENCH_NOCPU=1 timeout 600 python bench.py --steps 64 --warmup 8 > $O/${TAG}_bench_pf$pf.json 2> $O/${TAG}_bench_pf$pf.err; python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_pf$pf.json').read())
print('PF$pf value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'])
PY
done
