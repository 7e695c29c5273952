#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_num_owner -c 1 -f -o gpurun_out/owner_full \
    python tools/exp_spgemm.py 0.5 1 > gpurun_out/ncu_owner.log 2>&1
tail -2 gpurun_out/ncu_owner.log | cut -c1-200
