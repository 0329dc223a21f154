#!/bin/bash
# NOTE: this call ran the build with the 48 / 232 setmaxnreg split: every cc_token_kernel launch hung, pytest and the
# first bench hit their timeouts and the call was cut at its 2400 s limit (41 GPU-minutes, no data).  See r2/notes.md.
# round 2, call L: the whole GPU suite on the final token kernel, then the stall-gated L2 prefetch sweep
O=gpurun_out
mkdir -p $O
TAG=${TAG:-r2l}
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -q -m gpu -x > $O/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -6 $O/${TAG}_pytest.log
for pf in ${PFS:-0 4 8 12 0}; do
  CC_B200_TOKEN_PREFETCH=$pf CC_BENCH_NOCPU=1 timeout 300 python bench.py --steps 128 --warmup 8 > $O/${TAG}_pf${pf}_bench.json 2> $O/${TAG}_pf${pf}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_pf${pf}_bench.err; python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_pf${pf}_bench.json').read())
    print('PF$pf value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stream frac', d['step_stream']['frac_of_hbm_peak'], d['clocks'])
except Exception as e: print('no bench line', e)
PY
done
CC_B200_TOKEN_PREFETCH=8 timeout 200 python profiles/trace_token.py > $O/${TAG}_pf8_trace.txt 2>&1; grep -v Warn $O/${TAG}_pf8_trace.txt | grep -v nanmean | tail -34
