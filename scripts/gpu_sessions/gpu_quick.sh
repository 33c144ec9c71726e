#!/bin/bash
# quick loop: fused parity tests + timing of the fused kernel (+ optional ncu)
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or golden or crossings or ragged or deferred or checkpoint or energy" ) 2>&1 | tail -8 | tee gpurun_out/pytest_quick.log
timeout 300 python scripts/ab_kernels.py 128 64 0 fusedonly 2>&1 | tee gpurun_out/ab_quick.log
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block -s 3 -c 1 \
  -o gpurun_out/prof_axis_block_q -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_q.log 2>&1
fi
