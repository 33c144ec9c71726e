#!/bin/bash
# full default bench (with e2e phases and the CPU baseline) + launch list of the fused step
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 1500 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused2.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches2.log 2>&1
