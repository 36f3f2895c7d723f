#!/bin/bash
# round 2, call B: GPU tests on the single-pass assembly, A/B against the two-pass build, bench
set -u
TAG=${1:-r2b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 600 bash tools/ab_time.sh 1581 > /dev/null 2>&1; cp $OUT/ab_1581.log $OUT/${TAG}_ab_1581.log; cat $OUT/${TAG}_ab_1581.log
timeout 900 bash tools/ab_time.sh 3536 > /dev/null 2>&1; cp $OUT/ab_3536.log $OUT/${TAG}_ab_3536.log; cat $OUT/${TAG}_ab_3536.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
cut -c1-600 $OUT/${TAG}_bench_50M.json
