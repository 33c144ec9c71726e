#!/bin/bash
# ncu of the fused axis block + launch list of a fused step
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block -s 3 -c 1 \
  -o gpurun_out/prof_axis_block -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_block.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
