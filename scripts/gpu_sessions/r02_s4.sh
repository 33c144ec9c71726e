#!/bin/bash
# round 2, session 4 (2 GPUs): TMA alignment probe, full suite (user W on the hot kernels, TMA rows, slabs at world 2), TMA A/B
set -x
mkdir -p gpurun_out
for args in "3 8 4 1024 3" "3 8 4 1024 4" "4 8 4 12 3" "4 8 4 12 2" "2 4 32 1024 6" "2 4 32 1024 8"; do timeout 60 scripts/micro/tma_probe3 $args; done 2>&1 | grep -v "^+" | tee gpurun_out/r02_s4_tma_probe3.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r02_s4_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s4_pytest.log | head -40
for t in 0 1 0 1; do
timeout 300 python bench.py --no-e2e --no-cpu --no-secondary --steps 2 --warmup 3 --opt tma=$t 2>> gpurun_out/r02_s4_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tma=$t', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'])
" | tee -a gpurun_out/r02_s4_bench_tma_ab.txt
done
