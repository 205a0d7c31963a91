#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --maxfail 20 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
for wl in C4 C2 C3; do timeout 600 python bench.py --workload $wl --no-e2e --no-cpu-baseline > gpurun_out/bench5_$wl.json 2>> gpurun_out/bench.err; done
tail -4 gpurun_out/pytest.log; tail -3 gpurun_out/bench.err
# compute-sanitizer memcheck on the small smoke case (one tool per call)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py --smoke > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
tail -6 gpurun_out/memcheck.log
