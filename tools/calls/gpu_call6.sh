#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 600 > gpurun_out/pytest_mr.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mr.log
timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/scale_n2.json 2>> gpurun_out/scale.err
echo "n2 exit $?" >> gpurun_out/scale.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/ref_n2.json 2>> gpurun_out/scale.err
tail -5 gpurun_out/pytest_mr.log; tail -5 gpurun_out/scale.err; head -c 400 gpurun_out/scale_n2.json
