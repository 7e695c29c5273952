#!/bin/bash
# expand/sort/compress SpGEMM path: parity tests, then the configs[3] mult_ab leg with the phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_cuda_large.py tests/test_cuda_golden.py tests/test_cuda_property.py -x -q -m gpu -k "${K:-spgemm or multiply or mult_ab or item_item}" 2>&1 | tail -8
echo "== cfg3"; CSRK_TRACE=1 timeout 900 python tools/exp_cfg3_ab.py ${SCALE:-1.0} ${FRAC:-0.079} 2 2>&1 | grep -v "^\[csrk\]  *\(sort\|radix\|transpose\|scan\)" | tail -30
if [ -n "$NCU" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv python tools/exp_cfg3_ab.py ${SCALE:-1.0} ${FRAC:-0.079} 1 > gpurun_out/ncu_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_cfg3.csv 2>/dev/null | head -24
fi
