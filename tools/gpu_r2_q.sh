#!/bin/bash
# mask-guided fill pass: parity first (whole GPU suite with it on), then A/B timing at the bench size
set -u
TAG=${1:-r2_fillm}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
N=${2:-3536}
LOG=$OUT/${TAG}_ab_$N.log
echo "== register-staged fill (FCFV_FILL_MASKED=0)" | tee $LOG
FCFV_FILL_MASKED=0 timeout 300 python tools/quick_time.py $N 2>&1 | tee -a $LOG
echo "== mask-guided fill, default build" | tee -a $LOG
timeout 300 python tools/quick_time.py $N 2>&1 | tee -a $LOG
for so in fcfv_nme23_b200/variants/*.so; do
  [ -e "$so" ] || continue
  echo "== $so" | tee -a $LOG
  FCFV_LIB=$PWD/$so timeout 300 python tools/quick_time.py $N 2>&1 | tee -a $LOG
done
