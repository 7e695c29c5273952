#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== spgemm parity tests"; timeout 1200 python -m pytest tests/test_cuda_large.py tests/test_cuda_golden.py -x -q -m gpu -k "item or mult or fixed or spgemm or abt" 2>&1 | tail -4
echo "== fix threads"; CSRK_TRACE=0 timeout 600 python tools/exp_fix.py 2>&1 | tail -4
echo "== trace"; CSRK_TRACE=1 timeout 600 python tools/exp_spgemm.py 1.0 2 2>&1 | tail -40
