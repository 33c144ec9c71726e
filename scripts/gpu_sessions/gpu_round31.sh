#!/bin/bash
# step path without host synchronisation: parity + timing on a quiet and on a loaded host, against the previous build
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or golden or crossings or ragged or deferred or gauss or energy" ) 2>&1 | tail -3 | tee gpurun_out/pytest_quick.log
run() { # $1 = lib, $2 = label
  SPIC_B200_LIBRARY=$PWD/strugepic_b200/lib/$1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2 $1', round(d['ms_per_step'],1), 'kernels', round(sum(d['kernel_ms_per_step'].values()),1), 'launches', d['gpu_launches'])"
}
run libstrugepic_b200_sync.so quiet
run libstrugepic_b200.so quiet
pids=""
for i in $(seq 1 $(( $(nproc) * 2 ))); do python -c "while True: pass" & pids="$pids $!"; done
sleep 1
run libstrugepic_b200_sync.so loaded
run libstrugepic_b200.so loaded
kill $pids
