#!/bin/bash
# round 2, session 2c (2 GPUs): slab parity at world 2 (overlapped + serial exchange), 2-GPU drivers, field-only parity
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_drivers.py tests/test_parity_gpu.py -m gpu -q ) 2>&1 | tail -25 | tee gpurun_out/r02_s2c_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 2> gpurun_out/r02_s2c_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for s in d['secondary']: print(s.get('name'), s.get('value'), s.get('ms_per_step'), s.get('roofline',{}).get('frac'), s.get('error'))
" | tee gpurun_out/r02_s2c_bench.txt
