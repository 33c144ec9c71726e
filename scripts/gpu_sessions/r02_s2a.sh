#!/bin/bash
# round 2, session 2a (1 GPU): suite after the comm / gauss / bench rework, full bench line (checks + secondaries),
# low-ppc shape with the launch-per-sub-flow schedule (v3 stream kernels) against the fused block
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/r02_s2a_pytest.log
timeout 900 python bench.py > gpurun_out/r02_s2a_bench.json 2> gpurun_out/r02_s2a_bench.err
tail -c 600 gpurun_out/r02_s2a_bench.err; cat gpurun_out/r02_s2a_bench.json
timeout 600 python bench.py --cells 512 --ppc 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --no-fuse > gpurun_out/r02_s2a_bench512x8_nofuse.json 2> gpurun_out/r02_s2a_bench512x8_nofuse.err
tail -c 300 gpurun_out/r02_s2a_bench512x8_nofuse.err; cat gpurun_out/r02_s2a_bench512x8_nofuse.json
