#!/bin/bash
# round 2, session 10 (1 GPU): x-wall box (MABC + reflection) at 256^3 x 64 ppc: half-block schedule vs launch per sub-flow
set -x
mkdir -p gpurun_out
for f in "" "--no-fuse"; do
timeout 600 python bench.py --walls $f --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s10_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('walls $f', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['particles_conserved'], d['config']['particles'])
" | tee -a gpurun_out/r02_s10_bench_walls_ab.txt
done
tail -3 gpurun_out/r02_s10_bench.err
