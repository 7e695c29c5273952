#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
echo "== spgemm"; timeout 900 python tools/exp_spgemm.py 1.0 3 0.02 2>&1 | tail -8
echo "== spgemm cfg4 x0.1"; timeout 900 python tools/exp_spgemm.py 0.05 1 0.1 2>&1 | tail -4
