#!/bin/bash
# round-2 evidence: default bench + reference arm, launch list of the bench command, full captures of the two
# dominant kernels (k_spmv_slab, k_num_fixed) -- summaries go to profiles/ (tools/ncu_summary.py, launch_summary.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | head -c 600
echo "== launch list (bench, spmv + spgemm legs only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --skip-cfg0 --zipf-skew 0 --cfg3-scale 0 --skip-spgemm-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
echo "== ncu full: slab"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_slab -s 3 -c 1 -f -o gpurun_out/slab_full \
    python tools/exp_slab.py 1.0 16:1024:4096:2:2 > gpurun_out/ncu_slab.log 2>&1
tail -2 gpurun_out/ncu_slab.log
echo "== ncu full: fixed"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_num_fixed -s 1 -c 1 -f -o gpurun_out/fixed2_full \
    python tools/exp_spgemm.py 1.0 2 > gpurun_out/ncu_fx2.log 2>&1
tail -2 gpurun_out/ncu_fx2.log
