#!/bin/bash
# the driver's own launch at N=2 with default flags (all legs, e2e included), then the reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/s33_n2_default.json 2> gpurun_out/s33.err
echo "n2 default exit $?"
python -c "
import json; d=json.load(open('gpurun_out/s33_n2_default.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d.get('e2e'), d.get('gpu_launches'), d.get('cpu_baseline'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s33_n2_reference.json 2>> gpurun_out/s33.err
echo "n2 reference exit $?"; head -c 400 gpurun_out/s33_n2_reference.json; echo
tail -3 gpurun_out/s33.err
