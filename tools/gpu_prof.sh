#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_cuda_psf.py -x -q -k "powerlaw_many_slabs and f4-f4" > gpurun_out/racecheck.log 2>&1
grep -E "Error|Warning|hazard|access at" gpurun_out/racecheck.log | head -24
echo "== ncu full psf"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_psf_spmv -s 3 -c 1 -f -o gpurun_out/psf_full \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_psf.log 2>&1
tail -2 gpurun_out/ncu_psf.log | cut -c1-300
