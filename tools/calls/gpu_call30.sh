#!/bin/bash
# final artifacts of this session: GPU suite, smoke(), default bench line, ncu launch list, ncu --set full of the
# three C4 kernels and of the fused kernel at C2 / C3 shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest30.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest30.log; tail -3 gpurun_out/pytest30.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r01c_bench_C4.json 2> gpurun_out/bench30.err
echo "bench exit $?"; head -c 200 gpurun_out/r01c_bench_C4.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_bench_reference.json 2>> gpurun_out/bench30.err
echo "reference arm exit $?"; head -c 300 gpurun_out/r01c_bench_reference.json; echo
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
timeout 300 $CMD > gpurun_out/plain30.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_ncu_launches.csv $CMD > gpurun_out/ncu30_list.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi2E' -s 8 -c 1 -o gpurun_out/r01c_prof_fused $CMD > gpurun_out/ncu30_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_fwd_tma' -s 4 -c 1 -o gpurun_out/r01c_prof_fwd $CMD > gpurun_out/ncu30_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi1E' -s 4 -c 1 -o gpurun_out/r01c_prof_bwd $CMD > gpurun_out/ncu30_bwd.log 2>&1
for f in fused fwd bwd; do tail -n 1 gpurun_out/ncu30_$f.log; done
for wl in C2 C3; do
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_bwd_tma.*Li16ELi2E' -s 8 -c 1 -o gpurun_out/r01c_prof_fused_$wl $CMD > gpurun_out/ncu30_$wl.log 2>&1
  tail -n 1 gpurun_out/ncu30_$wl.log
done
