#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_psf3_spmv -s 2 -c 1 -f -o gpurun_out/psf3_full \
    python tools/exp_mode.py 3 > gpurun_out/ncu_psf3.log 2>&1
tail -2 gpurun_out/ncu_psf3.log | cut -c1-200
