#!/bin/bash
# N GPUs (8): multi-GPU tests on real devices, weak-scaling bench c4 (with the strong point, parity record, e2e, single-process
# multi-GPU handle), bench c5 (100 M elements), host-side timing of the multi-GPU handle
set -u
TAG=${1:-r2_8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multi_handle.py tests/test_multigpu_nccl.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_weak_${N}gpu.json 2> $OUT/${TAG}_bench_weak_${N}gpu.err; echo "bench c4 rc=$?"
grep "^\[bench" $OUT/${TAG}_bench_weak_${N}gpu.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --config c5 > $OUT/${TAG}_bench_c5_${N}gpu.json 2> $OUT/${TAG}_bench_c5_${N}gpu.err; echo "bench c5 rc=$?"
grep "^\[bench" $OUT/${TAG}_bench_c5_${N}gpu.err | tail -3
timeout 600 python tools/time_multi_handle.py 2000 $N > $OUT/${TAG}_multi_handle_time.log 2>&1; tail -3 $OUT/${TAG}_multi_handle_time.log
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_weak_${N}gpu.json", "$OUT/${TAG}_bench_c5_${N}gpu.json"):
    try:
        d = json.load(open(f))
        print(d["config"]["workload"][:60]); print(" value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
        print(" strong", d.get("strong")); print(" parity", d.get("multi_gpu_parity"))
        e = d.get("e2e") or {}; print(" e2e", e.get("value"), (e.get("mesh_resident") or {}).get("value"), "single", e.get("single_process_multi_gpu_handle"))
        print(" kernels", d["roofline"]["kernel_ms_per_step"])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
