#!/bin/bash
cd "$(dirname "$0")/.."
echo "== full"; CSRK_TRACE=1 timeout 600 python tools/exp_block.py 1 0 2>&1 | grep -E "spgemm: numeric|products \+ symbolic|^rank" | tail -4
echo "== blocks of 8"; timeout 600 python tools/exp_block.py 8 0,1,7 2>&1 | grep "^rank"
echo "== fixed test (one)"; timeout 1200 python -m pytest tests -m gpu -q -x -k "fixed_point_path and f8" 2>&1 | tail -3
