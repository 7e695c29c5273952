#!/bin/bash
# N-GPU bench (and, with TESTS=1, the multi-GPU parity tests) on one box: N=2 by default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
if [ "${TESTS:-0}" = "1" ]; then
  echo "== dist tests"; timeout 900 python -m pytest tests/test_cuda_dist.py -x -q -m gpu 2>&1 | tail -5
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 100 --warmup 5 ${EXTRA} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; grep -c "^{" gpurun_out/bench_n$N.json; tail -c 1500 gpurun_out/bench_n$N.err | tail -8
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}); print(d['roofline']['kernel'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])
print(d.get('collectives')); print({k:v for k,v in d.get('spgemm',{}).items() if k in ('value','ms','parity')}); print(d.get('cfg4'))
PY
