#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --maxfail 30 -x > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-other-workloads > gpurun_out/bench11.json 2> gpurun_out/bench11.err
echo "bench exit $?"; tail -3 gpurun_out/bench11.err
python -c "
import json; d=json.load(open('gpurun_out/bench11.json'))
print('value %.4g ms %.4f host_us %.1f'%(d['value'],d['ms_per_step'],d['host_us_per_step'])); print(d['e2e'])"
timeout 300 python tools/host_profile.py C1 2>&1 | head -3
