#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c1_reference_golden.py -x -q 2>&1 | tail -5
