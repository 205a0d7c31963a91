#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_one_launch.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -4
