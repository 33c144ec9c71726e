#!/bin/bash
# round 2, session 16 (1 GPU): PWL in-cell forms without range tests: suite + PWL throughput + issue-slot counters
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -6 | tee gpurun_out/r02_s16_pytest.log
for i in 1 2; do
timeout 600 python bench.py --interp pwl --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s16_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pwl', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['gauss_ok'])
" | tee -a gpurun_out/r02_s16_bench_pwl.txt
done
timeout 600 python bench.py --interp pwl --cells 512 --ppc 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s16_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pwl 512^3 x 8', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_ok'])
" | tee -a gpurun_out/r02_s16_bench_pwl.txt
tail -3 gpurun_out/r02_s16_bench.err
