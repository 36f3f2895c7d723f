#!/bin/bash
# new host-buffer paths (face lists, fcfv_compute_fcfv pipeline): their tests first, then the whole GPU suite and the bench
set -u
TAG=${1:-r2_v8}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_sparse_face_data.py tests/test_compute_fcfv.py tests/test_upload_cache.py -m gpu -q > $OUT/${TAG}_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -15 $OUT/${TAG}_pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 600 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
tail -5 $OUT/${TAG}_bench_50M.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
e = d["e2e"]
nel = 50013184
for tag, x in (("with_mesh", e), ("resident", e["mesh_resident"]), ("cache_off", e["cache_off"])):
    print("e2e", tag, round(x["value"], 1), "ms", round(x["ms_per_step"], 1), "up B/el", x["h2d_bytes_per_step"] / nel, "down", x["d2h_bytes_per_step"] / nel)
print(e.get("face_lists"), e.get("compute_fcfv_pipelined_calls"))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
