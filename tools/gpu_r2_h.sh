#!/bin/bash
set -u
TAG=${1:-r2h2}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_handle.py tests/test_multigpu_nccl.py tests/test_upload_cache.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 600 python tools/time_multi_handle.py 1000 $N > $OUT/${TAG}_multi_handle_time.log 2>&1; cat $OUT/${TAG}_multi_handle_time.log
timeout 600 python tools/time_multi_handle.py 2000 $N >> $OUT/${TAG}_multi_handle_time.log 2>&1; tail -4 $OUT/${TAG}_multi_handle_time.log
