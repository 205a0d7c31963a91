#!/bin/bash
# First GPU call: parity tests, variant sweep, bench, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; free -g | head -2 >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --maxfail 20 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python tools/sweep.py --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/sweep.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bwd -s 4 -c 1 -o gpurun_out/prof_bwd $CMD > gpurun_out/ncu_bwd.log 2>&1
timeout 300 $CMD > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 4 -c 1 -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu_fwd.log 2>&1
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/sweep.log; cat gpurun_out/bench.json | head -c 1500
