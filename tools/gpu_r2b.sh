#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_cuda_stream.py -x -q -m gpu 2>&1 | tail -5
echo "== exp_stream 1.0"; timeout 900 python tools/exp_stream.py 1.0 31,16 2>&1 | tail -8
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 2 -c 1 -f -o gpurun_out/stream_full \
    python tools/exp_stream.py 1.0 31 > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
