#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_num_fixed -c 1 -f -o gpurun_out/fixed_full \
    python tools/exp_spgemm.py 1.0 1 > gpurun_out/ncu_fx.log 2>&1; tail -1 gpurun_out/ncu_fx.log | cut -c1-80
