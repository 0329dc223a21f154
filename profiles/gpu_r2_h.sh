#!/bin/bash
# round 2, call H: A/B of token-kernel builds (profiles/libcc_<name>.so): token-kernel tests, bench, stage trace
O=gpurun_out
mkdir -p $O
TAG=${TAG:-r2h}
for name in "$@"; do
  export CC_B200_LIB=profiles/libcc_$name.so
  echo "=== build $name"
  timeout 420 python -m pytest tests/test_gpu_token_kernel.py tests/test_gpu_model.py -x -q -m gpu > $O/${TAG}_${name}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${TAG}_${name}_tests.log
  CC_BENCH_NOCPU=1 timeout 300 python bench.py --steps 64 --warmup 8 > $O/${TAG}_${name}_bench.json 2> $O/${TAG}_${name}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_${name}_bench.err; python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_${name}_bench.json').read())
    print('$name value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'])
except Exception as e: print('no bench line', e)
PY
  timeout 200 python profiles/trace_token.py > $O/${TAG}_${name}_trace.txt 2>&1; grep -v Warn $O/${TAG}_${name}_trace.txt | grep -v nanmean | tail -34
done
