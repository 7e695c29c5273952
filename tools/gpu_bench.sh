#!/bin/bash
# small-scale shake-out of every bench leg, then the default bench and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== small"; timeout 600 python bench.py --scale 0.05 --spgemm-scale 0.05 --cfg3-scale 0.01 --cfg3-products 1e7 --steps 20 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { tail -20 gpurun_out/bench_small.err; exit 1; }
echo "== full"; timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -12 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | head -c 9000
echo "== reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
