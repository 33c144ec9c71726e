#!/bin/bash
# round 2, session 18 (1 GPU): HEAD with the TMA-tiled curl sweeps on by default: suite, smoke, the default bench line
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -6 | tee gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
tail -c 300 gpurun_out/r02_final_bench.err; cut -c 1-300 gpurun_out/r02_final_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_256c64ppc.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_launches.log 2>&1
