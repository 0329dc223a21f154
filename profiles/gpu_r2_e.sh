#!/bin/bash
O=gpurun_out
for w in 8 16; do
  export CC_B200_LIB=profiles/libcc_w$w.so
  echo "=== warps $w"
  timeout 600 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu 2>&1 | tail -2
  CC_BENCH_NOCPU=1 timeout 600 python bench.py --steps 64 --warmup 8 > $O/r2p_bench_w$w.json 2> $O/r2p_bench_w$w.err; python - <<PY
import json
d=json.loads(open('$O/r2p_bench_w$w.json').read())
print('W$w value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'])
PY
  timeout 300 python profiles/trace_token.py > $O/r2p_trace_w$w.txt 2>&1; grep -A8 "stage  wait" $O/r2p_trace_w$w.txt | grep -v Warn | grep -v nanmean; grep "layer period" $O/r2p_trace_w$w.txt
done
export CC_B200_LIB=profiles/libcc_w8.so
python profiles/microtrace_token.py > $O/r2p_micro_w8.txt 2>&1; grep -A3 "stage 53" $O/r2p_micro_w8.txt
