#!/bin/bash
# density-loader twin test + compute-sanitizer (memcheck, racecheck) on the fused path with the new record layout
set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "loader" ) 2>&1 | tail -3 | tee gpurun_out/pytest_loader.log
timeout 80 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck.log
timeout 80 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck.log
