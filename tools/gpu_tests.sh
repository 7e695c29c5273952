#!/bin/bash
# the whole GPU suite, as the driver runs it, plus smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 3000 python -m pytest tests/ -q -m gpu 2>&1 | tail -15
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
