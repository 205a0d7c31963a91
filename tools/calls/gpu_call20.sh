#!/bin/bash
# N-GPU validation of the in-kernel tail exchange (peer mailboxes) + NCCL fallback, then the scaling points
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --no-other-workloads --no-e2e > gpurun_out/s20_n$n.json 2>> gpurun_out/s20.err
    echo "n$n exit $?"
    python -c "
import json; d=json.load(open('gpurun_out/s20_n$n.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d.get('exchange'), d['gpu_launches'], d['clocks'])"
  fi
done
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-other-workloads --no-e2e > gpurun_out/s20_n1.json 2>> gpurun_out/s20.err
python -c "
import json; d=json.load(open('gpurun_out/s20_n1.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['clocks'])"
tail -5 gpurun_out/s20.err
