#!/bin/bash
set -u
TAG=${1:-r2k}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python tools/quick_time.py 3536 > $OUT/${TAG}_quick_3536.log 2>&1; tail -2 $OUT/${TAG}_quick_3536.log
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > $OUT/${TAG}_bench_c5_1gpu.json 2> $OUT/${TAG}_bench_c5_1gpu.err; echo "c5 rc=$?"
tail -3 $OUT/${TAG}_bench_c5_1gpu.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_c5_1gpu.json"))
print(d["config"]["workload"]); print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"]); print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
