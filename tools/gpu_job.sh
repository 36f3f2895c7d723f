#!/bin/bash
# Standard GPU job (run through gpurun from the repo root):  tools/gpu_job.sh <tag> [bench|prof|all]
#   tests -> bench (50M default) -> ncu launch list + one --set full capture at S(1581).
# Numbers printed by runs under ncu are never bench values.
set -u
TAG=${1:-dev}
WHAT=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
if [ "$WHAT" = all ] || [ "$WHAT" = test ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
  cat $OUT/${TAG}_bench_50M.json
  timeout 300 python tools/quick_time.py 1581 > $OUT/${TAG}_quick_1581.log 2>&1; cat $OUT/${TAG}_quick_1581.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = prof ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --mesh-n 1581 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_asm_fill|k_asm_count|k_res_elem|k_local' -s 4 -c 4 \
      -o $OUT/${TAG}_full -f python bench.py --mesh-n 1581 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
if [ "$WHAT" = all ] || [ "$WHAT" = traffic ]; then
  # DRAM bytes of the hot kernels at the bench's own size (50M elements), one step
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'k_asm_fill|k_asm_count|k_res_elem|k_local|k_elem_coef|k_res_late' -s 7 -c 7 \
      -o $OUT/${TAG}_traffic50M -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
fi
