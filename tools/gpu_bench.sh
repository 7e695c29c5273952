#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json
  tail -4 gpurun_out/bench_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 100 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
  tail -6 gpurun_out/bench_n$N.err
fi
