#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --maxfail 30 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
