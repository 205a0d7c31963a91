#!/bin/bash
# same as gpu_call30.sh without the test suite; the .ncu-rep files are turned into raw CSV / details text on the box
# (five full captures exceed the 64 MiB that travel back)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r01c_bench_C4.json 2> gpurun_out/bench31.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_bench_reference.json 2>> gpurun_out/bench31.err
echo "reference arm exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads"
timeout 300 $CMD > gpurun_out/plain31.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_ncu_launches.csv $CMD > gpurun_out/ncu31_list.log 2>&1
echo "launch list exit $?"
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$rx" -s $skip -c 1 -o /tmp/$name "$@" > gpurun_out/ncu31_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r01c_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/r01c_ncu_${name}_details.txt 2>/dev/null
  tail -n 1 gpurun_out/ncu31_$name.log
}
cap fused 'k_bwd_tma.*Li16ELi2E' 8 $CMD
cp /tmp/fused.ncu-rep gpurun_out/r01c_prof_fused.ncu-rep
cap fwd 'k_fwd_tma' 4 $CMD
cap bwd 'k_bwd_tma.*Li16ELi1E' 4 $CMD
for wl in C2 C3; do
  cap fused_$wl 'k_bwd_tma.*Li16ELi2E' 8 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads
done
du -sh gpurun_out
