#!/bin/bash
# scaling points with the per-rank uncoupled step time next to the coupled one
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --no-other-workloads --no-e2e > gpurun_out/s26_n$n.json 2>> gpurun_out/s26.err
    echo "n$n exit $?"
    python -c "
import json; d=json.load(open('gpurun_out/s26_n$n.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), ['%.4f'%x for x in d.get('uncoupled_ms_per_rank',[])], d['clocks'])"
  fi
done
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-other-workloads --no-e2e > gpurun_out/s26_n1.json 2>> gpurun_out/s26.err
python -c "
import json; d=json.load(open('gpurun_out/s26_n1.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['clocks'])"
tail -3 gpurun_out/s26.err
