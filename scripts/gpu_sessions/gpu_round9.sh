#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block_persistent -s 3 -c 1 \
  -o gpurun_out/prof_axis_block_p -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_block_p.log 2>&1
ls -la gpurun_out
