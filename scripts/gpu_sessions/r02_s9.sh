#!/bin/bash
# round 2, session 9 (8 GPUs): slab parity at world 8 (eight cases in one rendezvous + one through pytest), N=8 weak scaling
set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  tests/mgpu_worker.py p8 pwl p8_thin p8_nofuse p8_tall pwl_tall p8_tall_serial pwl_serial_thin ) > gpurun_out/r02_s9_pytest_world8.log 2>&1
grep -E "parity ok|PARITY FAILED|rror" gpurun_out/r02_s9_pytest_world8.log | head -20
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "8-user_tall" ) >> gpurun_out/r02_s9_pytest_world8.log 2>&1
grep -E "passed|failed" gpurun_out/r02_s9_pytest_world8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_s9_bench_n8.json 2> gpurun_out/r02_s9_bench_n8.err
tail -c 400 gpurun_out/r02_s9_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s9_bench_n8.json").read().strip().splitlines()[-1])
print("N=8", d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["checks"]["gauss_drift_max"], d["checks"]["particles_conserved"], d["checks"]["gauss_ok"])
PY
