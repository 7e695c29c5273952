#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== large tests (fixed-point gate)"; timeout 900 python -m pytest tests/test_cuda_large.py -x -q -m gpu -k "signed" 2>&1 | grep -v "^$" | tail -30
for m in 2 4; do echo "== colmul $m"; CSRK_TRACE=1 timeout 600 python tools/exp_slab.py 1.0 16:1024:4096:2:2 1.0 $m 2>&1 | grep -v "sort:\|transpose" | tail -12; done
