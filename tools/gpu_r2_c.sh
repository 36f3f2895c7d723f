#!/bin/bash
# 1 GPU: upload-cache tests + C-ABI tests + a parity subset, then the default bench
set -u
TAG=${1:-r2c}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_upload_cache.py tests/test_cabi.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench_50M.err | tail -20
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("value", d["value"], "ms", d["ms_per_step"])
e = d["e2e"]
for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "upload_cache_hit_bytes_per_step", "mesh_resident", "cache_off", "fused_pass"):
    print(k, e.get(k))
print("cpu", d.get("cpu_baseline"))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
