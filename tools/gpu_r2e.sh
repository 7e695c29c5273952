#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_cuda_stream.py -x -q -m gpu 2>&1 | tail -3
echo "== exp_slab 1.0"; timeout 900 python tools/exp_slab.py 1.0 ${CFGS:-16:1024:4096:2:2} 2>&1 | tail -12
echo "== launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_slab.csv python tools/exp_slab.py 1.0 16:1024:4096:2:2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_slab.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[-14:]: print(r[ki][:60], r[vi])
PY
