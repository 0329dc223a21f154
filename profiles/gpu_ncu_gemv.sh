#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmsnorm_ffn_gate -s 40 -c 1 -f -o $O/r1_ffn_gate_full python profiles/run_gemv.py > $O/r1_ncu_gemv.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_residual -s 400 -c 1 -f -o $O/r1_w2_full python profiles/run_gemv.py >> $O/r1_ncu_gemv.log 2>&1; echo "rc=$?"
tail -3 $O/r1_ncu_gemv.log
