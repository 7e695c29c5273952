#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_cuda_stream.py -x -q -m gpu 2>&1 | tail -15
echo "== exp_slab 0.1"; timeout 300 python tools/exp_slab.py 0.1 16:512:4096 2>&1 | tail -4
echo "== exp_slab 1.0"; timeout 900 python tools/exp_slab.py 1.0 16:512:4096,16:512:8192,8:512:8192,12:512:4096,16:1024:4096,24:512:4096 2>&1 | tail -8
