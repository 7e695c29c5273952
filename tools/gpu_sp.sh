#!/bin/bash
# SpGEMM check: parity tests of the dense paths, then configs[2] timing with the phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== spgemm parity tests"; timeout 1200 python -m pytest tests/test_cuda_large.py tests/test_cuda_golden.py tests/test_cuda_property.py -x -q -m gpu -k "${K:-item_item or signed or multiply or mult_ab}" 2>&1 | tail -3
echo "== trace"; CSRK_TRACE=1 timeout 600 python tools/exp_spgemm.py 1.0 3 2>&1 | grep "numeric\|symbolic\|mult_abt" | tail -6
