#!/bin/bash
# round 2, session 2e (2 GPUs): split axis block after the double-push fix, TMA staging parity + A/B
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py -m gpu -q -k "slab or batch_shapes" ) > gpurun_out/r02_s2e_pytest.log 2>&1
grep -E "passed|failed|FAILED|parity ok|PARITY|rror" gpurun_out/r02_s2e_pytest.log | head -40
for t in 0 1 0 1; do
timeout 300 python bench.py --no-e2e --no-cpu --no-secondary --steps 2 --warmup 3 --opt tma=$t 2>> gpurun_out/r02_s2e_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tma=$t', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'])
" | tee -a gpurun_out/r02_s2e_bench_tma_ab.txt
done
