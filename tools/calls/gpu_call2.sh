#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --maxfail 20 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 900 python tools/sweep.py --out gpurun_out/sweep2.json > gpurun_out/sweep2.log 2>&1
echo "sweep exit $?" >> gpurun_out/sweep2.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -8 gpurun_out/pytest.log; tail -2 gpurun_out/sweep2.log; head -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
