#!/bin/bash
mkdir -p gpurun_out
for wl in C2 C3; do
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
  timeout 300 $CMD > gpurun_out/plain_$wl.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi2E' -s 4 -c 1 -o gpurun_out/prof18_fused_$wl $CMD > gpurun_out/ncu18_$wl.log 2>&1
  tail -n 1 gpurun_out/ncu18_$wl.log
done
