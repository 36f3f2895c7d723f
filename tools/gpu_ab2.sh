#!/bin/bash
# A/B of the default library and every build under fcfv_nme23_b200/variants/ at S(N): step and per-kernel times
set -u
TAG=${1:-ab}; N=${2:-3536}
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/${TAG}_ab_$N.log
echo "== default build" | tee $LOG
timeout 300 python tools/quick_time.py $N 2>&1 | grep "^step\|k_local=" | tee -a $LOG
for so in fcfv_nme23_b200/variants/*.so; do
  [ -e "$so" ] || continue
  echo "== $so" | tee -a $LOG
  FCFV_LIB=$PWD/$so timeout 300 python tools/quick_time.py $N 2>&1 | grep "^step\|k_local=" | tee -a $LOG
done
