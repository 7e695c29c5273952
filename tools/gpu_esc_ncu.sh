#!/bin/bash
# full ncu capture of one kernel of the expand/sort/compress path on a scaled-down configs[3] block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN:-k_esc_sortmerge} -c 1 -f -o gpurun_out/${OUT:-esc_sortmerge} python tools/exp_cfg3_ab.py ${SCALE:-0.2} ${FRAC:-0.03} 1 > gpurun_out/ncu_esc.log 2>&1
tail -3 gpurun_out/ncu_esc.log
