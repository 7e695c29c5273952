#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== psf tests"; timeout 600 python -m pytest tests/test_cuda_psf.py -x -q -k celltile 2>&1 | tail -25
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_cuda_psf.py -x -q -k "celltile and powerlaw_many_slabs and f4-f4" 2>&1 | tail -6
echo "== split exp"; timeout 600 python tools/exp_split.py 2>&1 | tail -8
