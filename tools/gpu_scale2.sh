#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
[ "$N" = "2" ] && { echo "== dist test"; timeout 600 python -m pytest tests/test_cuda_dist.py -x -q 2>&1 | tail -4; }
for ch in 1 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 100 --warmup 5 --spgemm-scale 0 --chunks $ch > gpurun_out/bench_n${N}_c$ch.json 2> gpurun_out/bench_n${N}_c$ch.err
echo "chunks=$ch rc=$?"; tail -c 300 gpurun_out/bench_n${N}_c$ch.err | tail -1
done
