#!/bin/bash
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
echo "== debug"; timeout 200 python profiles/debug_prefill.py > $O/${TAG}_dbg_prefill.log 2>&1; echo "rc=$?"; tail -2 $O/${TAG}_dbg_prefill.log
echo "== pytest prefill"; timeout 600 python -m pytest tests/test_gpu_prefill.py -q -m gpu > $O/${TAG}_pytest_prefill.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest_prefill.log
echo "== prefill timing"; timeout 300 python profiles/run_prefill.py > $O/${TAG}_prefill_time.txt 2>&1; cat $O/${TAG}_prefill_time.txt
echo "== ncu full: prefill kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_kernel -s 1 -c 1 \
    -f -o $O/${TAG}_prefill_full python profiles/run_prefill.py --once --shape llama8k > $O/${TAG}_ncu_prefill.log 2>&1; echo "ncu prefill rc=$?"
