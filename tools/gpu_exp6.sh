#!/bin/bash
cd "$(dirname "$0")/.."
echo "== gpu tests (large)"; timeout 900 python -m pytest tests/test_cuda_large.py tests/test_cuda_golden.py -m gpu -q -x 2>&1 | tail -5
echo "== own"; timeout 900 python tools/exp_own.py 2>&1 | tail -5
