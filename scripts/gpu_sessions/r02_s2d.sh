#!/bin/bash
# round 2, session 2d (2 GPUs): slab parity at world 2 incl. tall slabs (split axis block, overlapped exchange), field-only
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py -m gpu -q -k "slab or field_only" ) > gpurun_out/r02_s2d_pytest.log 2>&1
grep -v "^$" gpurun_out/r02_s2d_pytest.log | grep -E "passed|failed|FAILED|parity|PARITY|rror|assert" | head -60
timeout 300 python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 2> gpurun_out/r02_s2d_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for s in d['secondary']: print(s.get('name'), s.get('value'), s.get('ms_per_step'), s.get('roofline',{}).get('frac'), s.get('error'))
" | tee gpurun_out/r02_s2d_bench.txt
