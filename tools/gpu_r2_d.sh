#!/bin/bash
# 1 GPU: full GPU test suite, A/B of the variants at S(3536), default bench
set -u
TAG=${1:-r2d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 900 bash tools/ab_time.sh 3536 > /dev/null 2>&1; cp $OUT/ab_3536.log $OUT/${TAG}_ab_3536.log; cat $OUT/${TAG}_ab_3536.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
tail -5 $OUT/${TAG}_bench_50M.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("value", d["value"], "ms", d["ms_per_step"])
e = d["e2e"]
for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "upload_cache_hit_bytes_per_step", "mesh_resident", "cache_off", "fused_pass"):
    print(k, e.get(k))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
