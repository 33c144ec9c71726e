#!/bin/bash
# round 2, session 1: first GPU run of k_axis_block_s (parity, then A/B timing), full GPU suite, low-ppc shape
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/r02_s1_pytest.log
timeout 600 python tests/tools/check_block_stream.py 128 2>&1 | tail -20 | tee gpurun_out/r02_s1_check_block_stream.log
for bs in 0 1; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --opt block_stream=$bs > gpurun_out/r02_s1_bench256_bs$bs.json 2> gpurun_out/r02_s1_bench256_bs$bs.err
  tail -c 300 gpurun_out/r02_s1_bench256_bs$bs.err; cat gpurun_out/r02_s1_bench256_bs$bs.json
  timeout 600 python bench.py --cells 512 --ppc 8 --steps 2 --warmup 3 --no-e2e --no-cpu --opt block_stream=$bs > gpurun_out/r02_s1_bench512x8_bs$bs.json 2> gpurun_out/r02_s1_bench512x8_bs$bs.err
  tail -c 300 gpurun_out/r02_s1_bench512x8_bs$bs.err; cat gpurun_out/r02_s1_bench512x8_bs$bs.json
done
ls -la gpurun_out
