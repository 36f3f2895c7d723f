#!/bin/bash
# what the driver runs at round end, on one GPU: GPU test suite, smoke, both bench arms
set -u
TAG=${1:-r2_v6}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > $OUT/${TAG}_bench_50M.json 2> $OUT/${TAG}_bench_50M.err; echo "bench rc=$?"
python - <<PY
import json
r = json.load(open("$OUT/${TAG}_bench_reference.json"))
d = json.load(open("$OUT/${TAG}_bench_50M.json"))
print("reference", r["value"], r["unit"])
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "clocks", d["clocks"])
e = d["e2e"]
print("e2e", e["value"], "resident", e["mesh_resident"]["value"], "ratio e2e/reference", e["value"] / r["value"])
print("roofline frac", d["roofline"]["frac"], "pipeline", d["roofline"]["pipeline"]["frac"], "assembly", d["roofline"]["assembly_phase"]["frac"], "residual", d["roofline"]["residual_phase"]["frac"])
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
