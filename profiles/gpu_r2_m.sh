#!/bin/bash
# round 2, call M: final bench line of the token kernel + the whole GPU suite, every step under a hard timeout
O=gpurun_out
mkdir -p $O
TAG=${TAG:-r2m}
CC_BENCH_NOCPU=1 timeout 150 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
timeout 420 python -m pytest tests -q -m gpu -x --timeout 150 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${TAG}_pytest.log
