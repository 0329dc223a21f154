#!/bin/bash
# round 2, call G: the whole GPU suite, the bench lines of both arms, the token kernel's stage trace, and the ncu
# evidence (launch list of the bench command + one --set full capture of cc_token_kernel)
O=gpurun_out
mkdir -p $O
TAG=${1:-r2g}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_smi.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > $O/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -6 $O/${TAG}_pytest.log
echo "== bench (ours)"; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "rc=$?"; tail -2 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "rc=$?"; cat $O/${TAG}_bench_ref.json
echo "== stage trace"; timeout 300 python profiles/trace_token.py > $O/${TAG}_trace.txt 2>&1; tail -40 $O/${TAG}_trace.txt | grep -v Warn | grep -v nanmean
echo "== ncu launch list"
CC_BENCH_NOCPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_list.log 2>&1; echo "rc=$?"
grep -c cc_token_kernel $O/${TAG}_launches.csv
echo "== ncu full (token kernel)"
CC_BENCH_NOCPU=1 CC_BENCH_NCU=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:cc_token_kernel -c 1 -o $O/${TAG}_tok_full -f python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_full.log 2>&1; echo "rc=$?"
ls -la $O | tail -12
