#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== split exp"; timeout 600 python tools/exp_split.py 2>&1 | tail -7
echo "== spgemm"; timeout 900 python tools/exp_spgemm.py 1.0 2 0.02 2>&1 | tail -8
echo "== ncu launches spgemm"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_spgemm.csv python tools/exp_spgemm.py 0.5 1 0.01 > gpurun_out/ncu_spgemm.log 2>&1
tail -3 gpurun_out/ncu_spgemm.log
