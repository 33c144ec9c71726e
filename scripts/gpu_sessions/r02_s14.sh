#!/bin/bash
# round 2, session 14 (1 GPU): pair kernel with conflict-free record stores: suite, low-count shapes, ncu counters
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -6 | tee gpurun_out/r02_s14_pytest.log
for cfg in "512 8" "384 16"; do set -- $cfg
timeout 600 python bench.py --cells $1 --ppc $2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s14_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1^3 x $2 ppc', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['particles_conserved'])
" | tee -a gpurun_out/r02_s14_bench_lowppc.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block_pair -s 3 -c 1 \
  -o gpurun_out/r02_prof_axis_block_pair_v2 -f python bench.py --cells 256 --ppc 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_pair_v2.log 2>&1
