#!/bin/bash
# round 2, session 7 (1 GPU): low-count push_V_E kernel, adapter test, suite, low-ppc shapes, headline unchanged
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r02_s7_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s7_pytest.log | head -40
for cfg in "512 8" "384 16" "256 32"; do set -- $cfg
timeout 600 python bench.py --cells $1 --ppc $2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s7_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1^3 x $2 ppc', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['particles_conserved'])
" | tee -a gpurun_out/r02_s7_bench_lowppc.txt
done
timeout 600 python bench.py --cells 256 --ppc 32 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --opt pair_kernel=1 2>> gpurun_out/r02_s7_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('256^3 x 32 ppc pair_kernel=1', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
" | tee -a gpurun_out/r02_s7_bench_lowppc.txt
tail -5 gpurun_out/r02_s7_bench.err
