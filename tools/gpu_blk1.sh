#!/bin/bash
cd "$(dirname "$0")/.."
echo "== spgemm tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "spgemm or item_item or mult or golden or virtual or sharded" 2>&1 | tail -3
echo "== blocks of 8"; timeout 600 python tools/exp_block.py 8 0,1,7 2>&1 | grep "^rank"
echo "== blocks of 2"; timeout 600 python tools/exp_block.py 2 0,1 2>&1 | grep "^rank"
echo "== full"; CSRK_TRACE=1 timeout 600 python tools/exp_block.py 1 0 2>&1 | tail -14
