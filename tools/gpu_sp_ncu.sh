#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN:-k_num_fixed} -s 1 -c 1 -f -o gpurun_out/${OUT:-fixed2_full} python tools/exp_spgemm.py 1.0 2 > gpurun_out/ncu_fx2.log 2>&1
tail -3 gpurun_out/ncu_fx2.log
