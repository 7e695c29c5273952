#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== spgemm"; timeout 900 python tools/exp_spgemm.py 1.0 3 2>&1 | tail -4
echo "== ncu launches spgemm cfg3 full"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_spgemm3.csv python tools/exp_spgemm.py 1.0 1 > gpurun_out/ncu_spgemm3.log 2>&1
tail -2 gpurun_out/ncu_spgemm3.log
