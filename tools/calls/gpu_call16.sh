#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 800 > gpurun_out/pytest_mr.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mr.log
tail -6 gpurun_out/pytest_mr.log
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-e2e --no-other-workloads > gpurun_out/s16_n1.json 2> gpurun_out/s16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-e2e --no-other-workloads > gpurun_out/s16_n2.json 2>> gpurun_out/s16.err
echo "n2 exit $?"
for f in s16_n1 s16_n2; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], 'value %.4g ms %.4f two-pass %.4g' % (d['value'], d['ms_per_step'], d['two_pass']['value']), d.get('exchange'), d['kernel_level_step_ms']['one_pass'])"; done
grep -v "^\*\|OMP_NUM\|Warn\|warn\|^$" gpurun_out/s16.err | tail -5
