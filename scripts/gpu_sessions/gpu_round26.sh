#!/bin/bash
# ncu --set full with source of the fused axis block (128^3 x 64 ppc) for a per-line instruction / stall breakdown
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block -s 3 -c 1 \
  -o gpurun_out/prof_axis_block_r28 -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_r28.log 2>&1
tail -3 gpurun_out/ncu_r28.log; ls -la gpurun_out/
