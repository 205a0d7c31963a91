#!/bin/bash
# BASELINE config 5 across the 8 GPUs of one box: global batch 64..1024, text 32 / 256
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 tools/c5_sweep_dist.py > gpurun_out/c5_dist.log 2>&1
echo "exit $?"; grep "^{" gpurun_out/c5_dist.log | tail -12; grep -i "error\|Traceback" gpurun_out/c5_dist.log | head -5
