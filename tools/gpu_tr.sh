#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/exp_tr.py 0.2 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tr.csv python tools/exp_tr.py 0.2 > gpurun_out/ncu_tr.log 2>&1
