#!/bin/bash
# round-1 final verification of HEAD: all GPU tests, smoke, both bench arms, ncu launch list, ncu --set full of the two top kernels
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 1500 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_continue -s 2 -c 1 \
  -o gpurun_out/prof_axis_continue -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_cont.log 2>&1
