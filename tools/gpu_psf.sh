#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== psf tests"; timeout 600 python -m pytest tests/test_cuda_psf.py -x -q 2>&1 | tail -25
echo "== racecheck (small psf case)"
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=k_psf_spmv python -m pytest tests/test_cuda_psf.py -x -q -k "powerlaw_many_slabs and f4-f4" 2>&1 | tail -8
echo "== memcheck (small psf case)"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_cuda_psf.py -x -q -k "powerlaw_many_slabs and f4-f4" 2>&1 | tail -8
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -3 gpurun_out/bench.err
