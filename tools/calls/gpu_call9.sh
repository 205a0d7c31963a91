#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --maxfail 30 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err
echo "bench exit $?"; head -c 300 gpurun_out/bench9.json
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(fwd|bwd|epilogue|modality)' -c 200 --csv --log-file gpurun_out/launches9.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi2E' -s 4 -c 1 -o gpurun_out/prof9_fused $CMD > gpurun_out/ncu_fused.log 2>&1
timeout 300 $CMD > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_fwd_tma' -s 4 -c 1 -o gpurun_out/prof9_fwd $CMD > gpurun_out/ncu_fwd.log 2>&1
timeout 300 $CMD > gpurun_out/plain4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi1E' -s 4 -c 1 -o gpurun_out/prof9_bwd $CMD > gpurun_out/ncu_bwd.log 2>&1
for f in fused fwd bwd; do tail -n 2 gpurun_out/ncu_$f.log; done
