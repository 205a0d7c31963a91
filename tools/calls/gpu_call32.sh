#!/bin/bash
# device-side epochs everywhere: multirank parity incl. CUDA-graph replay of the sharded step, N=2 bench with trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_one_launch.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
MAFED_B200_DIST=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/multirank_check.py 2>&1 | grep -i "graph\|fail\|error" | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --no-other-workloads --no-e2e > gpurun_out/s32_n2.json 2>> gpurun_out/s32.err
echo "n2 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/s32_n2.json')); print(d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), ['%.4f'%x for x in d.get('uncoupled_ms_per_rank',[])], d.get('exchange_trace_us'), d['clocks'])"
