#!/bin/bash
# multi-GPU call: NCCL paths (handle-internal and one-process-per-GPU), bench with the in-run parity record
set -u
TAG=${1:-r2m}; N=${2:-4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multi_handle.py tests/test_multigpu_nccl.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${N}gpu.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "parity", d.get("multi_gpu_parity"), "e2e", d.get("e2e", {}).get("value"))
except Exception as ex:
    print("bench output unreadable:", ex); print(open("$OUT/${TAG}_bench_${N}gpu.err").read()[-3000:])
PY
