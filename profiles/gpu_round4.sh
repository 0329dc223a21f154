#!/bin/bash
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
