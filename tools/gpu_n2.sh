#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
for mode in fused nofused; do
flag=""; [ "$mode" = "nofused" ] && flag="--no-fused"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 100 --warmup 5 --spgemm-scale 0 $flag > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
tail -c 400 gpurun_out/bench_n${N}_$mode.err | tail -2
done
