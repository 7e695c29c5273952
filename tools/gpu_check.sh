#!/bin/bash
# GPU-box script: smoke, parity tests, bench, ncu launch list + one full capture of the SpMV kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 --spgemm-scale ${SPGEMM_SCALE:-0.25} 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== ncu full (SpMV tile kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tile -s 3 -c 2 -f -o gpurun_out/spmv_full \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
