#!/bin/bash
# end-of-session check on a 2-GPU box: whole GPU suite (multi-GPU tests included: late rank, graph replay), smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
