#!/bin/bash
# Times every A/B build under fcfv_nme23_b200/variants/ (and the default library) with tools/quick_time.py.
N=${1:-1581}
OUT=gpurun_out
mkdir -p $OUT
echo "== default" | tee $OUT/ab_$N.log
python tools/quick_time.py $N 2>&1 | tee -a $OUT/ab_$N.log
for so in fcfv_nme23_b200/variants/*.so; do
  [ -e "$so" ] || continue
  echo "== $so" | tee -a $OUT/ab_$N.log
  FCFV_LIB=$PWD/$so timeout 300 python tools/quick_time.py $N 2>&1 | tee -a $OUT/ab_$N.log
done
