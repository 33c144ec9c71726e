#!/bin/bash
# round 2, session 6 (1 GPU): two-cells-per-batch kernel (k_axis_block_pair): suite, then 512^3 x 8 ppc with and without it
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_s6_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s6_pytest.log | head -40
for pk in 1 0; do
timeout 600 python bench.py --cells 512 --ppc 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --opt pair_kernel=$pk 2>> gpurun_out/r02_s6_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pair_kernel=$pk', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['particles_conserved'])
" | tee -a gpurun_out/r02_s6_bench_pair_ab.txt
done
timeout 600 python bench.py --cells 384 --ppc 16 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --opt pair_kernel=1 2>> gpurun_out/r02_s6_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('384^3 x 16 ppc pair_kernel=1', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
" | tee -a gpurun_out/r02_s6_bench_pair_ab.txt
timeout 600 python bench.py --cells 384 --ppc 16 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --opt pair_kernel=0 2>> gpurun_out/r02_s6_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('384^3 x 16 ppc pair_kernel=0', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
" | tee -a gpurun_out/r02_s6_bench_pair_ab.txt
tail -5 gpurun_out/r02_s6_bench.err
