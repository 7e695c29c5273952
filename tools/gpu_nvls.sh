#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
echo "== dist test"; timeout 600 python -m pytest tests/test_cuda_dist.py -q -x 2>&1 | tail -5
for mode in on off; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 100 --warmup 5 --nvls $mode --spgemm-scale 0 > gpurun_out/bench_n${N}_nvls_$mode.json 2> gpurun_out/bench_n${N}_nvls_$mode.err
  echo "nvls=$mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_nvls_$mode.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("collectives"), d["gpu_launches"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_n${N}_nvls_$mode.err").read()[-1500:])
PY
done
