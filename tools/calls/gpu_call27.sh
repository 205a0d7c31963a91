#!/bin/bash
# LL-word exchange: multirank parity (peer + nccl), then coupled / uncoupled step times and the exchange trace
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --no-other-workloads --no-e2e > gpurun_out/s27_n$n.json 2>> gpurun_out/s27.err
    echo "n$n exit $?"
    python -c "
import json; d=json.load(open('gpurun_out/s27_n$n.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), ['%.4f'%x for x in d.get('uncoupled_ms_per_rank',[])], d.get('exchange_trace_us'), d['clocks'])"
  fi
done
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-other-workloads --no-e2e > gpurun_out/s27_n1.json 2>> gpurun_out/s27.err
python -c "
import json; d=json.load(open('gpurun_out/s27_n1.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['clocks'])"
tail -3 gpurun_out/s27.err
