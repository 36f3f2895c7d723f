#!/bin/bash
set -u
TAG=${1:-ab}; N=${2:-3536}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 bash tools/ab_time.sh $N > /dev/null 2>&1; cp $OUT/ab_$N.log $OUT/${TAG}_ab_$N.log; cat $OUT/${TAG}_ab_$N.log
