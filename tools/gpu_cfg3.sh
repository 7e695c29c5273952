#!/bin/bash
# configs[3] mult_ab: phase trace at full scale, then an ncu launch list of one product
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CSRK_TRACE=1 timeout 900 python tools/exp_cfg3_ab.py ${SCALE:-1.0} ${FRAC:-0.079} 2 2>&1 | grep -v "^\[csrk\] \(radix\|transpose\|scan\)" | tail -40
if [ -n "$NCU" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv python tools/exp_cfg3_ab.py ${SCALE:-1.0} ${FRAC:-0.079} 1 > gpurun_out/ncu_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_cfg3.csv 2>/dev/null | head -30
fi
