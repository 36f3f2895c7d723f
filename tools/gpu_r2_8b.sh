#!/bin/bash
# 8 GPUs, short: the weak-scaling bench line (with the strong point, the in-run parity record, e2e and the single-process multi-GPU handle)
set -u
TAG=${1:-r2_v9_8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_weak_${N}gpu.json 2> $OUT/${TAG}_bench_weak_${N}gpu.err; echo "bench rc=$?"
grep "^\[bench" $OUT/${TAG}_bench_weak_${N}gpu.err | tail -3
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_weak_${N}gpu.json"))
    print(d["config"]["workload"][:60]); print(" value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
    print(" strong", d.get("strong")); print(" parity", d.get("multi_gpu_parity"))
    e = d.get("e2e") or {}; print(" e2e", e.get("value"), (e.get("mesh_resident") or {}).get("value"), "check", e.get("result_check"), "single", e.get("single_process_multi_gpu_handle"))
except Exception as ex:
    print("unreadable:", ex); print(open("$OUT/${TAG}_bench_weak_${N}gpu.err").read()[-2000:])
PY
