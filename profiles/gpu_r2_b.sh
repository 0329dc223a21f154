#!/bin/bash
# round 2, call B: first light of the whole-token kernel
O=gpurun_out
mkdir -p $O
echo "== attention stage traces"; timeout 600 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu -k "attention_stage" > $O/r2b_attn.log 2>&1; echo "rc=$?"; tail -15 $O/r2b_attn.log
echo "== whole token"; timeout 600 python -m pytest tests/test_gpu_token_kernel.py -x -q -m gpu -k "not attention_stage" > $O/r2b_tok.log 2>&1; echo "rc=$?"; tail -30 $O/r2b_tok.log
