#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python scripts/ab_kernels.py 128 64 0 2>&1 | tee gpurun_out/ab_p8.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2000 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_theta_axis_v2 -s 6 -c 1 \
  -o gpurun_out/prof_theta_axis_v2b -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
