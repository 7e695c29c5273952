#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== transpose"; CSRK_TRACE=1 timeout 600 python tools/exp_tr.py 0.2 2 2>&1 | tail -24
