#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_cuda_stream.py -x -q -m gpu 2>&1 | tail -15
echo "== dropin tests"; timeout 900 python -m pytest tests/test_cuda_dropin.py -x -q -m gpu 2>&1 | tail -15
echo "== exp_stream 0.1"; timeout 300 python tools/exp_stream.py 0.1 31,16 2>&1 | tail -8
echo "== exp_stream 1.0"; CSRK_TRACE=0 timeout 900 python tools/exp_stream.py 1.0 31,23,16 2>&1 | tail -8
