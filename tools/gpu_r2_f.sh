#!/bin/bash
# N GPUs: full GPU suite, quick timing, then the N-rank bench on a SMALL mesh with a short timeout (stall hunt), then at full size
set -u
TAG=${1:-r2f2}; N=${2:-2}; SMALL=${3:-1581}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 300 python tools/quick_time.py 3536 > $OUT/${TAG}_quick_3536.log 2>&1; tail -2 $OUT/${TAG}_quick_3536.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --mesh-n $SMALL > $OUT/${TAG}_bench_small_${N}gpu.json 2> $OUT/${TAG}_bench_small_${N}gpu.err; rc=$?; echo "small bench rc=$rc"
grep "^\[bench" $OUT/${TAG}_bench_small_${N}gpu.err | tail -20
if [ $rc -eq 0 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
  grep "^\[bench" $OUT/${TAG}_bench_${N}gpu.err | tail -20
  python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_${N}gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
print("strong", d.get("strong")); print("parity", d.get("multi_gpu_parity"))
e = d["e2e"]; print("e2e", e["value"], e["mesh_resident"]["value"], "single", e.get("single_process_multi_gpu_handle"))
print("kernels", d["roofline"]["kernel_ms_per_step"])
PY
else
  tail -30 $OUT/${TAG}_bench_small_${N}gpu.err
fi
