#!/bin/bash
# end-of-session check on a fresh box: what the driver runs (GPU suite, smoke, default bench, reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r01d_bench_C4.json 2> gpurun_out/bench37.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r01d_bench_C4.json')); print('value %.4g ms %.4f frac %.3f launches %d'%(d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']), d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
