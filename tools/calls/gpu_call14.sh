#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --maxfail 30 -x > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
echo "bench exit $?"; tail -3 gpurun_out/bench14.err
python -c "
import json; d=json.load(open('gpurun_out/bench14.json'))
print('value %.4g ms %.4f host_us %.1f'%(d['value'],d['ms_per_step'],d['host_us_per_step']), d['kernel_level_step_ms'], 'fused ms', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])
for k,v in d['other_workloads'].items(): print(k, 'value %.4g ms %.4f frac %.3f'%(v['value'], v['ms_per_step'], v['frac']))"
