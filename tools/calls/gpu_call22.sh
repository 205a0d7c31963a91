#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/sweep_ring.py --workloads C4,C2,C3 --iters 100 --repeats 3 --out gpurun_out/sweep_ring.json > gpurun_out/sweep_ring.log 2>&1
echo "exit $?"; wc -l gpurun_out/sweep_ring.log; grep -c error gpurun_out/sweep_ring.log
