#!/bin/bash
# one-launch step: new tests, whole GPU suite, A/B against separate launches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_one_launch.py -x -q -m gpu 2>&1 | tail -15
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python tools/ab_tail.py > gpurun_out/ab_tail.log 2>&1; cat gpurun_out/ab_tail.log
