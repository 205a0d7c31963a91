#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_reference_golden.py -x -q --durations=5 2>&1 | tail -12
