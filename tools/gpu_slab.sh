#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_cuda_stream.py -x -q -m gpu 2>&1 | tail -15
echo "== exp_slab 1.0"; timeout 900 python tools/exp_slab.py 1.0 ${CFGS:-16:1024:4096} 2>&1 | tail -8
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_slab -s 2 -c 1 -f -o gpurun_out/slab_full \
    python tools/exp_slab.py 1.0 ${NCU_CFG:-16:1024:4096} > gpurun_out/ncu_slab.log 2>&1
tail -3 gpurun_out/ncu_slab.log
