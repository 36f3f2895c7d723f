#!/bin/bash
# 1 GPU: full GPU suite, default bench, then the ncu passes (launch list of the bench's own command at its default size,
# --set full of the hot kernels at S(1581), DRAM traffic at 50 M elements).  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r2_v4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
timeout 300 python tools/quick_time.py 3536 > $OUT/${TAG}_quick_3536.log 2>&1; tail -2 $OUT/${TAG}_quick_3536.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_50M.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_asm_fill|k_asm_count|k_res_elem|k_local|k_res_late' -s 5 -c 5 \
    -o $OUT/${TAG}_full -f python bench.py --mesh-n 1581 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_asm_fill|k_asm_count|k_res_elem|k_local|k_elem_coef|k_res_late' -s 5 -c 5 \
    -o $OUT/${TAG}_traffic50M -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
e = d["e2e"]
for k in ("value", "ms_per_step", "h2d_bytes_per_step", "mesh_resident"):
    print(k, e.get(k))
print("cpu", d.get("cpu_baseline", {}).get("value"))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
