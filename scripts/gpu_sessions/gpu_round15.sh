#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/ab_kernels.py 128 64 0 fusedonly 2>&1 | tee gpurun_out/ab_p8.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.err; cat gpurun_out/bench.json
