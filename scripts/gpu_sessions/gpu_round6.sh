#!/bin/bash
# first run of the fused axis block: debug cases, sanitizer, parity tests, A/B timing, bench
set -x
mkdir -p gpurun_out
timeout 300 python tests/tools/debug_fused.py 2>&1 | tee gpurun_out/debug_fused.log | tail -60
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path[:0]=['.','oracle','tests']
import sys; sys.path.insert(0, "tests/tools"); import debug_fused as d
d.case((6,5,4), 40, 0.08, 0, 4, 1, 1)
d.case((6,5,4), 40, 0.3, 0, 4, 1, 1)
d.case((6,5,4), 40, 0.08, 1, 2, 1, 1)
" 2>&1 | tail -25 | tee gpurun_out/sanitizer.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/ab_kernels.py 128 64 0 2>&1 | tee gpurun_out/ab_p8.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.err; cat gpurun_out/bench.json
ls -la gpurun_out
