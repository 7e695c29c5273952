#!/bin/bash
# round-2 final evidence (second half): launch lists of the bench command and of the configs[3] mult_ab leg, full
# captures of the fixed-point SpGEMM kernel (lean and side-list variants) and of the three expand/sort/compress kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== launch list (bench, spmv + spgemm legs only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --skip-cfg0 --zipf-skew 0 --cfg3-scale 0 --skip-spgemm-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
echo "== launch list (configs[3] mult_ab)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv \
    python tools/exp_cfg3_ab.py 1.0 0.079 1 > gpurun_out/ncu_cfg3.log 2>&1
tail -1 gpurun_out/ncu_cfg3.log
echo "== ncu full: fixed (raw ratings: lean variant)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_num_fixed -s 1 -c 1 -f -o gpurun_out/fixed3_full \
    python tools/exp_spgemm_norm.py 1.0 2 raw > gpurun_out/ncu_fx3.log 2>&1
tail -1 gpurun_out/ncu_fx3.log
echo "== ncu full: fixed (mean-centred: side-list variant)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_num_fixed -s 1 -c 1 -f -o gpurun_out/fixed3c_full \
    python tools/exp_spgemm_norm.py 1.0 2 center > gpurun_out/ncu_fx3c.log 2>&1
tail -1 gpurun_out/ncu_fx3c.log
for k in k_esc_sortmerge k_esc_scatter k_esc_count; do
  echo "== ncu full: $k (configs[3] x0.2, 30 000 rows)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${k}_full \
      python tools/exp_cfg3_ab.py 0.2 0.03 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
