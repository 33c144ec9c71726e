#!/bin/bash
# round 2, session 8 (4 GPUs): slab parity at world 2 and 4 at HEAD (10 cases each), 2-GPU drivers, N=4 weak scaling with checks
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_drivers.py -m gpu -q ) > gpurun_out/r02_s8_pytest_mgpu.log 2>&1
grep -E "passed|failed|FAILED|PARITY FAILED|rror" gpurun_out/r02_s8_pytest_mgpu.log | head -30
grep -c "multi-gpu parity ok" gpurun_out/r02_s8_pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_s8_bench_n4.json 2> gpurun_out/r02_s8_bench_n4.err
tail -c 400 gpurun_out/r02_s8_bench_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s8_bench_n4.json").read().strip().splitlines()[-1])
print("N=4", d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["checks"]["gauss_drift_max"], d["checks"]["particles_conserved"], d["checks"]["gauss_ok"])
PY
