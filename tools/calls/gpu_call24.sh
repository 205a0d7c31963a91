#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/ab_fixup.py > gpurun_out/ab_fixup.log 2>&1; cat gpurun_out/ab_fixup.log
