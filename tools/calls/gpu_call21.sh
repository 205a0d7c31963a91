#!/bin/bash
# after the one-launch step: new tests, default bench (all legs), PCIe probe, ncu launch list + full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_one_launch.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/pcie_probe.py > gpurun_out/pcie_probe.json 2>&1; cat gpurun_out/pcie_probe.json
timeout 900 python bench.py > gpurun_out/r01b_bench_C4.json 2> gpurun_out/r01b_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r01b_bench_C4.json')); print('value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['roofline']['frac'], d['e2e'], d['clocks'])"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
timeout 300 $CMD > gpurun_out/plain21.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_ncu_launches.csv $CMD > gpurun_out/ncu21a.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi2E' -s 8 -c 1 -o gpurun_out/r01b_prof_fused $CMD > gpurun_out/ncu21b.log 2>&1
echo "full capture exit $?"; tail -n 2 gpurun_out/ncu21b.log
