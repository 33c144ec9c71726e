#!/bin/bash
# round 2, final verification of HEAD (1 GPU): all GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# ncu --set full of the kernels that changed this round
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/r02_final_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/r02_final_bench_reference_arm.err
cat gpurun_out/r02_final_bench_reference_arm.json | cut -c 1-400
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
tail -c 500 gpurun_out/r02_final_bench.err; cut -c 1-600 gpurun_out/r02_final_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_256c64ppc.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block -s 3 -c 1 \
  -o gpurun_out/r02_prof_axis_block -f python bench.py --cells 128 --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_axis_block_pair -s 3 -c 1 \
  -o gpurun_out/r02_prof_axis_block_pair -f python bench.py --cells 256 --ppc 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push_v_e_quad -s 3 -c 1 \
  -o gpurun_out/r02_prof_push_quad -f python bench.py --cells 256 --ppc 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/r02_ncu_quad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_curl -s 20 -c 2 \
  -o gpurun_out/r02_prof_curl -f python scripts/bench_field_only.py > gpurun_out/r02_ncu_curl.log 2>&1
( time timeout 600 drivers/bin/energy_conservation drivers/decks/energy_64.input nsteps=10001 print_every=500 ) > gpurun_out/r02_driver_energy_64_1e4_steps.txt 2>&1
tail -4 gpurun_out/r02_driver_energy_64_1e4_steps.txt
for tool in memcheck racecheck; do timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_small.py 2>&1 | grep -v "Host Frame" | tail -12; done | tee gpurun_out/r02_sanitizer.txt
ls -la gpurun_out | grep r02_prof
