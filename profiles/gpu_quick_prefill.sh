#!/bin/bash
# quick check of a prefill-kernel change: diagnostics + llama8k timing (+ optional ncu)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 200 python profiles/debug_prefill.py > $O/${TAG}_dbg_prefill.log 2>&1; echo "debug rc=$?"; tail -1 $O/${TAG}_dbg_prefill.log
timeout 300 python profiles/run_prefill.py --shape llama8k qwen16k > $O/${TAG}_prefill_time.txt 2>&1; grep -v "torch" $O/${TAG}_prefill_time.txt
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_kernel -s 1 -c 1 \
    -f -o $O/${TAG}_prefill_full python profiles/run_prefill.py --once --shape llama8k > $O/${TAG}_ncu_prefill.log 2>&1; echo "ncu prefill rc=$?"
fi
