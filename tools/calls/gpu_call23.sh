#!/bin/bash
mkdir -p gpurun_out
P="C4:mse:0,0,0,0;C4:mse:2,4,2,3;C4:mse:2,8,2,3;C4:mse:1,4,8,2;C4:mse:1,16,4,3;C4:mse:2,4,4,2;C4:mse:2,8,4,2"
P="$P;C4:cosine:0,0,0,0;C4:cosine:2,4,4,2;C4:cosine:1,16,4,4;C4:cosine:2,4,2,4;C4:cosine:2,4,2,3;C4:cosine:2,8,4,2"
P="$P;C2:mse:0,0,0,0;C2:mse:1,16,8,6;C2:mse:1,8,8,6;C2:mse:2,8,8,2;C2:mse:2,4,4,6;C2:mse:2,4,8,3;C2:mse:2,8,8,3;C2:mse:1,16,8,4"
P="$P;C2:cosine:0,0,0,0;C2:cosine:2,4,8,3;C2:cosine:2,4,8,2;C2:cosine:2,8,8,2;C2:cosine:1,16,8,6"
P="$P;C3:mse:0,0,0,0;C3:mse:1,8,16,2;C3:mse:2,8,4,3;C3:mse:2,4,4,3;C3:mse:2,4,8,2;C3:mse:2,8,8,2;C3:mse:1,16,8,4"
P="$P;C1:mse:0,0,0,0;C1:mse:2,4,4,3;C1:mse:1,16,8,4;C1:mse:2,8,4,3;C1:mse:2,4,4,2"
timeout 1500 python tools/sweep_ring.py --workloads C4,C2,C3,C1 --iters 200 --repeats 8 --points "$P" --out gpurun_out/sweep_ring2.json > gpurun_out/sweep_ring2.log 2>&1
echo "exit $?"; wc -l gpurun_out/sweep_ring2.log; grep -c error gpurun_out/sweep_ring2.log
