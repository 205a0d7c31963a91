#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 800 > gpurun_out/pytest_mr8.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mr8.log
tail -5 gpurun_out/pytest_mr8.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-other-workloads > gpurun_out/scale12_n$n.json 2>> gpurun_out/scale12.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --no-other-workloads > gpurun_out/scale12_n$n.json 2>> gpurun_out/scale12.err
    fi
    echo "n$n exit $?"
    python -c "
import json; d=json.load(open('gpurun_out/scale12_n$n.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d.get('exchange'), 'e2e', d.get('e2e',{}).get('value'), d['clocks'])"
  fi
done
