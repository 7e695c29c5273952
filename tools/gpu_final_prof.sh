#!/bin/bash
# Evidence run: launch list of the bench command + full captures of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tile -s 3 -c 1 -f -o gpurun_out/spmv_tile_full \
    python bench.py --steps 3 --warmup 3 --spgemm-scale 0 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-80
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_num_fixed -c 1 -f -o gpurun_out/fixed_full \
    python tools/exp_spgemm.py 1.0 1 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-80
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tr.csv \
    python tools/exp_tr.py 0.2 > /dev/null 2>&1; tail -2 gpurun_out/launches_tr.csv | cut -c1-100
