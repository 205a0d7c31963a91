#!/bin/bash
# refresh of the variants table (cosine / fp32 / fp16 / base / 410M, both kernel families) and of the C5 sweep on one GPU
mkdir -p gpurun_out
timeout 600 python tools/variants_bench.py > gpurun_out/variants34.log 2>&1; echo "variants exit $?"; tail -3 gpurun_out/variants34.log
timeout 900 python tools/c5_sweep.py > gpurun_out/c5_34.log 2>&1; echo "c5 exit $?"; tail -3 gpurun_out/c5_34.log
