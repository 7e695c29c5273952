#!/bin/bash
cd "$(dirname "$0")/.."
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== transpose"; timeout 600 python tools/exp_tr.py 0.2 2>&1 | tail -12
echo "== spgemm"; timeout 900 python tools/exp_spgemm.py 1.0 4 2>&1 | tail -5
