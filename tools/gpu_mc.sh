#!/bin/bash
cd "$(dirname "$0")/.."
N=${N:-2}
for v in 0 1 2 3 4 5; do
CSRK_MC_VARIANT=$v NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exp_mc.py 32000000 2>&1 | grep -E "variant|nccl|correct|Error|error" | tail -4
done
