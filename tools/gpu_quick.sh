#!/bin/bash
# quick correctness + timing check of the assembly: a few GPU parity tests, then tools/quick_time.py
set -u
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_golden_reference.py tests/test_gpu_parity.py -m gpu -x -q -k "${2:-H1 or LowRes or S6 or device_row or 50M or config3}" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
for n in 1581 3536; do timeout 300 python tools/quick_time.py $n > $OUT/${TAG}_quick_$n.log 2>&1; cat $OUT/${TAG}_quick_$n.log; done
