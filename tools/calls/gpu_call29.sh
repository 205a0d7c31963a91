#!/bin/bash
# 8-GPU box: parity of the LL-word exchange at N=8 (peer path), then the scaling points with trace
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
MAFED_B200_DIST=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/multirank_check.py > gpurun_out/mr29.log 2>&1
echo "multirank exit $?"; grep -c OK gpurun_out/mr29.log; grep -i "fail\|error\|exchange path" gpurun_out/mr29.log | sort | uniq -c | head -5
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --no-other-workloads --no-e2e > gpurun_out/s29_n$n.json 2>> gpurun_out/s29.err
    echo "n$n exit $?"
    python -c "
import json; d=json.load(open('gpurun_out/s29_n$n.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), ['%.4f'%x for x in d.get('uncoupled_ms_per_rank',[])], d.get('exchange_trace_us'), d['clocks'])"
  fi
done
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-other-workloads --no-e2e > gpurun_out/s29_n1.json 2>> gpurun_out/s29.err
python -c "
import json; d=json.load(open('gpurun_out/s29_n1.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['clocks'])"
