#!/bin/bash
# round 2, last run of HEAD (1 GPU): suite, smoke, the default bench line (refreshes profiles/r02_final_bench.json), ncu of the
# PWL block kernel
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -6 | tee gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
tail -c 300 gpurun_out/r02_final_bench.err; cut -c 1-300 gpurun_out/r02_final_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block -s 3 -c 1 \
  -o gpurun_out/r02_prof_axis_block_pwl -f python bench.py --cells 128 --interp pwl --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_block_pwl.log 2>&1
