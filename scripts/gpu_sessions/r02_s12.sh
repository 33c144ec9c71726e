#!/bin/bash
# round 2, session 12 (1 GPU): PWL block kernel at 2 / 3 / 4 resident blocks per SM (side-by-side builds, SPIC_PWL_BLOCKS)
set -x
mkdir -p gpurun_out
for lib in libstrugepic_b200_pwl2.so libstrugepic_b200.so libstrugepic_b200_pwl4.so; do
SPIC_B200_LIBRARY=$PWD/strugepic_b200/lib/$lib timeout 600 python bench.py --interp pwl --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s12_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['gauss_ok'])
" | tee -a gpurun_out/r02_s12_bench_pwl_blocks.txt
done
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "batch_shapes or fused_axis_block or golden" 2>&1 | tail -3
tail -3 gpurun_out/r02_s12_bench.err
