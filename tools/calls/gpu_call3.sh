#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/sweep_step.py --out gpurun_out/sweep3.json > gpurun_out/sweep3.log 2>&1
echo "sweep exit $?" >> gpurun_out/sweep3.log
tail -3 gpurun_out/sweep3.log
