#!/bin/bash
# secondary measurements: field_only 256^3 and the PWL variant of the headline workload
set -x
mkdir -p gpurun_out
timeout 300 python scripts/bench_field_only.py 256 200 2>&1 | tail -1 | tee gpurun_out/bench_field_only.json
timeout 600 python bench.py --interp pwl --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_pwl.json 2> gpurun_out/bench_pwl.err
tail -c 600 gpurun_out/bench_pwl.err; cat gpurun_out/bench_pwl.json
