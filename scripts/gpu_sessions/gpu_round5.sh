#!/bin/bash
# verification + profile pass of HEAD: GPU parity tests, both bench arms, ncu launch list, ncu --set full of the two particle kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv | tee gpurun_out/smi.txt
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.err; cat gpurun_out/bench.json
python scripts/ab_kernels.py 128 64 0 2>&1 | tee gpurun_out/ab_p8.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_theta_axis_v2 -s 6 -c 1 \
  -o gpurun_out/prof_theta_axis_v2 -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_axis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push_v_e -s 2 -c 1 \
  -o gpurun_out/prof_push_v_e -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_pushve.log 2>&1
ls -la gpurun_out
