#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --maxfail 30 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench.err
tail -30 gpurun_out/pytest.log; tail -3 gpurun_out/bench.err; head -c 250 gpurun_out/bench7.json
