#!/bin/bash
# batched tail-chunk loads (short rows): parity, then A/B against the previous build on the same box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python tools/ab_lib.py gpurun_ab/libmafed_distill_old.so 3 > gpurun_out/ab_lib.log 2>&1; cat gpurun_out/ab_lib.log
