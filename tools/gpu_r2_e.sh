#!/bin/bash
# 1 GPU: full GPU test suite, quick timing at S(3536), default bench
set -u
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 300 python tools/quick_time.py 3536 > $OUT/${TAG}_quick_3536.log 2>&1; cat $OUT/${TAG}_quick_3536.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
tail -3 $OUT/${TAG}_bench_50M.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
e = d["e2e"]
for k in ("value", "ms_per_step", "h2d_bytes_per_step", "mesh_resident"):
    print(k, e.get(k))
print("kernels", d["roofline"]["kernel_ms_per_step"])
print("roofline", {k: v for k, v in d["roofline"].items() if k in ("achieved", "frac", "kernel_ms")}, d["roofline"]["pipeline"], d["roofline"].get("assembly_phase"), d["roofline"].get("residual_phase"))
PY
