#!/bin/bash
# round 2, session 15 (2 GPUs): HEAD after the last kernel touches (leaver guard in the interior part, offset table in
# k_push_v_e_quad): whole suite incl. world-2 slabs, low-count shape, the full default bench line
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r02_s15_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s15_pytest.log | head -20
timeout 600 python bench.py --cells 512 --ppc 8 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>> gpurun_out/r02_s15_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('512^3 x 8 ppc', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['checks']['gauss_drift_max'], d['checks']['particles_conserved'])
" | tee -a gpurun_out/r02_s15_bench_lowppc.txt
timeout 900 python bench.py > gpurun_out/r02_s15_bench.json 2>> gpurun_out/r02_s15_bench.err
cut -c 1-300 gpurun_out/r02_s15_bench.json
tail -3 gpurun_out/r02_s15_bench.err
