#!/bin/bash
# fixed-point SpGEMM with equilibration + side list: parity tests, then configs[2] raw / centred / unit-normalised
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_cuda_large.py tests/test_cuda_golden.py tests/test_cuda_property.py tests/test_cuda_dropin.py -x -q -m gpu -k "${K:-spgemm or multiply or mult_ab or item_item or fixed}" 2>&1 | tail -8
echo "== configs[2]"; CSRK_TRACE=1 timeout 900 python tools/exp_spgemm_norm.py ${SCALE:-1.0} 2 2>&1 | grep "mult_abt\|numeric\|symbolic" | tail -24
