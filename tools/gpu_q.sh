#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== spmv tests"; timeout 900 python -m pytest tests -m gpu -x -q -k "spmv or mult_vec or golden" 2>&1 | tail -2
echo "== bench"; timeout 1200 python bench.py --spgemm-scale 0 2> gpurun_out/bench_q.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'])"
