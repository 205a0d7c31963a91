#!/bin/bash
# where does the cosine one-pass kernel lose time at 1.5 KB rows (base shape)?  one ncu --set full capture each, cosine and mse
mkdir -p gpurun_out
for loss in cosine mse; do
  CMD="python tools/sweep_ring.py --workloads C2 --losses $loss --points C2:$loss:0,0,0,0 --iters 5 --repeats 1"
  rx='k_bwd_tma.*Li1ELi16ELi2E'; [ $loss = mse ] && rx='k_bwd_tma.*Li0ELi16ELi2E'
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$rx" -s 4 -c 1 -o /tmp/c2_$loss $CMD > gpurun_out/ncu38_$loss.log 2>&1
  ncu -i /tmp/c2_$loss.ncu-rep --page details > gpurun_out/c2_${loss}_details.txt 2>/dev/null
  ncu -i /tmp/c2_$loss.ncu-rep --page raw --csv > gpurun_out/c2_${loss}_raw.csv 2>/dev/null
  tail -n 1 gpurun_out/ncu38_$loss.log
done
