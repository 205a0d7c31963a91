#!/bin/bash
# One gpurun call that produces the ncu evidence of a round (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r02'
# 1. plain run of the short bench command (must exit 0), 2. the same command under ncu for the launch list
# (gpu__time_duration.sum per launch: cold-cache and serialised -- compare SHARES), 3. `ncu --set full` of the fused
# kernel at the headline shape and of the cosine kernel at the base shape.  Outputs go to gpurun_out/<tag>_*.
set -u
tag=${1:-r02}
short="python bench.py --steps 3 --warmup 3 --no-c5 --no-other-workloads --no-e2e --no-cpu-baseline"
$short > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_ncu_launches.csv $short > gpurun_out/${tag}_ncu_launches.out 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bwd_tma -s 8 -c 2 -f -o gpurun_out/${tag}_ncu_fused_C4 $short > gpurun_out/${tag}_ncu_fused.out 2>&1
echo "full capture rc=$?"
cos="python tools/variants_bench.py --only cosine_base"
$cos > gpurun_out/${tag}_variants_plain.json 2> gpurun_out/${tag}_variants_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_bwd_tma -s 4 -c 1 -f -o gpurun_out/${tag}_ncu_fused_C2_cosine $cos > gpurun_out/${tag}_ncu_cos.out 2>&1
echo "cosine capture rc=$?"
ls -la gpurun_out/ | grep ${tag}_ncu
