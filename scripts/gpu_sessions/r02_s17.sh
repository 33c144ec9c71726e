#!/bin/bash
# round 2, session 17 (1 GPU): TMA-tiled curl sweeps (option curl_tma): parity, then A/B inside the PIC step (256^3: 3 + 4
# sweeps per step) and ncu of both kernels
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "curl_sweeps or field_only or every_subflow" ) 2>&1 | tail -15 | tee gpurun_out/r02_s17_pytest.log
for t in 0 1 0 1; do
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --opt curl_tma=$t 2>> gpurun_out/r02_s17_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('curl_tma=$t', d['value'], d['ms_per_step'], 'curl ms/step', d['kernel_ms_per_step']['curl'], d['checks']['gauss_ok'])
" | tee -a gpurun_out/r02_s17_bench_curl_tma_ab.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_curl -s 8 -c 4 \
  -o gpurun_out/r02_prof_curl_tma -f python bench.py --cells 256 --ppc 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary --opt curl_tma=1 > gpurun_out/r02_ncu_curl_tma.log 2>&1
tail -3 gpurun_out/r02_s17_bench.err
