#!/bin/bash
TAG=${1:-r1c}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu (model + parity)"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
echo "== gemv microbench"; timeout 300 python profiles/run_gemv.py > $O/${TAG}_gemv.txt 2>&1; head -6 $O/${TAG}_gemv.txt
echo "== bench"; timeout 900 python bench.py --steps 64 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/${TAG}_bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['us_per_launch'])"
for w in 2 8; do CC_TP_FAKE=$w timeout 300 python bench.py --steps 32 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['parallelism'], d['value'], d['ms_per_step'])"; done
