#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg4.csv python tools/exp_spgemm.py 0.02 1 0.05 > gpurun_out/ncu_cfg4.log 2>&1
tail -3 gpurun_out/ncu_cfg4.log
