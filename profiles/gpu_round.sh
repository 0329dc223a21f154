#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench command and
# ncu --set full captures of the fused decode kernel and the dominant GEMV.  Outputs -> gpurun_out/.
#   gpurun --timeout 1500 -- bash profiles/gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench.json
echo "== K1 timing"; timeout 300 python profiles/run_k1.py --time --steps 30 > $O/${TAG}_k1_time.txt 2>&1; cat $O/${TAG}_k1_time.txt
echo "== ncu launch list (bench)"
CC_BENCH_NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full: decode kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cc_decode_kernel -s 64 -c 3 \
    -f -o $O/${TAG}_k1_full python profiles/run_k1.py --steps 4 > $O/${TAG}_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
echo "== ncu full: gemv kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmsnorm_ffn_gate -s 4 -c 2 \
    -f -o $O/${TAG}_ffn_gate_full python profiles/run_gemv.py > $O/${TAG}_ncu_gemv.log 2>&1; echo "ncu gemv rc=$?"
ls -la $O
