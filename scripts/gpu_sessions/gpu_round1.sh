#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list, ncu --set full of the dominant kernel.
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_theta_axis_binned -s 6 -c 3 \
  -o gpurun_out/prof_theta_axis -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push_v_e_binned -s 2 -c 1 \
  -o gpurun_out/prof_push_v_e -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
python scripts/quick_perf.py 128 64 0 > gpurun_out/quick_p8.log 2>&1
python scripts/quick_perf.py 128 64 1 > gpurun_out/quick_pwl.log 2>&1
ls -la gpurun_out
