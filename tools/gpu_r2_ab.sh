#!/bin/bash
# A/B of the builds under fcfv_nme23_b200/variants: parity subset with each variant, then timings at S(3536)
set -u
TAG=${1:-r2ab}
OUT=gpurun_out; mkdir -p $OUT
for so in fcfv_nme23_b200/variants/*.so; do
  echo "== parity with $so" | tee -a $OUT/${TAG}_pytest.log
  FCFV_LIB=$PWD/$so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_reference.py -m gpu -x -q -k "local_assembly or device_pipeline or strips or large_mesh" >> $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?" | tee -a $OUT/${TAG}_pytest.log
done
tail -12 $OUT/${TAG}_pytest.log
timeout 900 bash tools/ab_time.sh 3536 > /dev/null 2>&1; cp $OUT/ab_3536.log $OUT/${TAG}_ab_3536.log; grep "^==\|^step\|k_local=" $OUT/${TAG}_ab_3536.log
