#!/bin/bash
# full GPU test suite, then the multi-GPU script (NCCL tests + N-rank bench with strong / parity / single-process records)
set -u
TAG=${1:-r2n}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $OUT/${TAG}_pytest_all.log
tail -6 $OUT/${TAG}_pytest_all.log
bash tools/gpu_r2_multi.sh $TAG $N
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_${N}gpu.json"))
print("strong", d.get("strong")); print("single", d.get("e2e", {}).get("single_process_multi_gpu_handle"))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
