#!/bin/bash
# round 2, call A: GPU test suite, compute-sanitizer (memcheck + racecheck) on the small cases, baseline bench
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > $OUT/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2a_pytest.log
tail -5 $OUT/r2a_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_cases.py > $OUT/r2a_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/r2a_memcheck.log
tail -4 $OUT/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_cases.py > $OUT/r2a_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/r2a_racecheck.log
tail -4 $OUT/r2a_racecheck.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r2a_bench_50M.json 2> $OUT/r2a_bench_50M.err; echo "bench rc=$?"
cat $OUT/r2a_bench_50M.json | cut -c1-1500
