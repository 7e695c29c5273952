#!/bin/bash
cd "$(dirname "$0")/.."
N=${N:-8}
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/exp_block_dist.py 2>&1 | grep -E "^rank|rank 0 rep|rank $((N-1)) rep|csrk\] (spgemm|transpose: enter)" | tail -90
