#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --maxfail 20 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 900 python tools/sweep_step.py --extra --out gpurun_out/sweep4.json > gpurun_out/sweep4.log 2>&1
echo "sweep exit $?" >> gpurun_out/sweep4.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
for wl in C2 C3; do timeout 600 python bench.py --workload $wl --no-e2e > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err; done
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(fwd|bwd|epilogue)' -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li8ELi2E' -s 4 -c 1 -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_fused.log 2>&1
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/sweep4.log; head -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ncu_fused.log
