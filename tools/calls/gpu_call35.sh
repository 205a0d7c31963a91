#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_intervals.py C2 C4 > gpurun_out/step_intervals.log 2>&1; cat gpurun_out/step_intervals.log
