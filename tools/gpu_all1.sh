#!/bin/bash
# 1-GPU round check: smoke, all gpu tests, default bench, reference arm (short)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"; timeout 1200 python bench.py 2> gpurun_out/bench_n1.err > gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-1500
