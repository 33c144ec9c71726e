#!/bin/bash
# round 2, session 11 (2 GPUs): whole suite after the Theta_E deferral rework (incl. world-2 slabs), wall box again
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r02_s11_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s11_pytest.log | head -30
timeout 600 python bench.py --walls --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s11_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('walls', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['gauss_ok'], d['checks']['particles_conserved'])
" | tee -a gpurun_out/r02_s11_bench_walls.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s11_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('periodic', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['gauss_ok'])
" | tee -a gpurun_out/r02_s11_bench_walls.txt
tail -3 gpurun_out/r02_s11_bench.err
