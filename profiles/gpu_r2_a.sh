#!/bin/bash
# round 2, call A: K1 phase breakdown + compute-sanitizer over K1 and the prefill kernel (item 9)
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt
echo "== K1 phases config 2 (8 heads x 1024)"; timeout 120 python profiles/trace_k1.py heavy_hitter 1024 > $O/r2_k1_phases_hh1024.txt 2>&1; tail -16 $O/r2_k1_phases_hh1024.txt
echo "== K1 phases l2 512"; timeout 120 python profiles/trace_k1.py l2 512 > $O/r2_k1_phases_l2_512.txt 2>&1; tail -16 $O/r2_k1_phases_l2_512.txt
echo "== K1 phases recent_global 64"; timeout 120 python profiles/trace_k1.py recent_global 64 > $O/r2_k1_phases_rg64.txt 2>&1; tail -16 $O/r2_k1_phases_rg64.txt
for pol in heavy_hitter l2 recent_global; do
  echo "== memcheck K1 $pol"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python profiles/run_k1.py --strategy $pol --layers 2 --steps 3 --L 256 > $O/r2_san_memcheck_k1_$pol.txt 2>&1; tail -4 $O/r2_san_memcheck_k1_$pol.txt
  echo "== racecheck K1 $pol"; timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python profiles/run_k1.py --strategy $pol --layers 2 --steps 3 --L 256 > $O/r2_san_racecheck_k1_$pol.txt 2>&1; tail -4 $O/r2_san_racecheck_k1_$pol.txt
done
echo "== memcheck prefill"; timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python profiles/run_prefill.py --once --shape llama4k > $O/r2_san_memcheck_prefill.txt 2>&1; tail -4 $O/r2_san_memcheck_prefill.txt
echo "== synccheck K1"; timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python profiles/run_k1.py --strategy heavy_hitter --layers 2 --steps 3 --L 256 > $O/r2_san_synccheck_k1_hh.txt 2>&1; tail -4 $O/r2_san_synccheck_k1_hh.txt
