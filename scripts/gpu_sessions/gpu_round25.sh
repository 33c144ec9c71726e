#!/bin/bash
# validation of the drivers, the density loader and the user-W slot; regression check of the headline kernel time
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_drivers.py tests/test_user_w.py -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_new.log
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_drivers.py --deselect tests/test_user_w.py ) 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
( time timeout 300 drivers/bin/energy_conservation drivers/decks/energy_64.input nsteps=300 ) 2>&1 | tail -12 | tee gpurun_out/driver_energy_64.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -c 600 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
