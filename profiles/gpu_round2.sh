#!/bin/bash
# GPU tests + prefill kernel timing + ncu captures of the prefill kernel.  gpurun --timeout 1200 -- bash profiles/gpu_round2.sh
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest_gpu.log
echo "== prefill timing"; timeout 300 python profiles/run_prefill.py > $O/${TAG}_prefill_time.txt 2>&1; cat $O/${TAG}_prefill_time.txt
echo "== ncu full: prefill kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_kernel -s 1 -c 1 \
    -f -o $O/${TAG}_prefill_full python profiles/run_prefill.py --once --shape llama8k > $O/${TAG}_ncu_prefill.log 2>&1; echo "ncu prefill rc=$?"
echo "== gemv microbench"; timeout 300 python profiles/run_gemv.py > $O/${TAG}_gemv.txt 2>&1; cat $O/${TAG}_gemv.txt
ls -la $O | tail -8
