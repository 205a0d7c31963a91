#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_adaptive_weights.py -m gpu -q --timeout 600 > gpurun_out/pytest_mr.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mr.log
tail -15 gpurun_out/pytest_mr.log
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-e2e > gpurun_out/scale8_n1.json 2> gpurun_out/scale8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-e2e > gpurun_out/scale8_n2_peer.json 2>> gpurun_out/scale8.err
echo "n2 peer exit $?" >> gpurun_out/scale8.err
MAFED_B200_DIST=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-e2e > gpurun_out/scale8_n2_nccl.json 2>> gpurun_out/scale8.err
echo "n2 nccl exit $?" >> gpurun_out/scale8.err
grep -v "^\*\|OMP_NUM\|Warn\|warn" gpurun_out/scale8.err | tail -8
for f in scale8_n1 scale8_n2_peer scale8_n2_nccl; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], 'value %.4g ms %.4f two-pass %.4g' % (d['value'], d['ms_per_step'], d['two_pass']['value']), d['gpu_launches'])"; done
