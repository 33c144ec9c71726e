#!/bin/bash
# round 2, session 19 (2 GPUs): slabs at world 2 with the TMA curl sweeps on by default + N=2 bench
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
  tests/mgpu_worker.py p8 pwl p8_thin p8_nofuse p8_tall pwl_tall p8_tall_serial pwl_serial_thin user_tall user_nofuse ) > gpurun_out/r02_s19_world2.log 2>&1
grep -E "parity ok|PARITY FAILED|rror" gpurun_out/r02_s19_world2.log | cut -c 1-160
timeout 600 python -m pytest tests/test_drivers.py -m gpu -q -k "two_gpus" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_s19_bench_n2.json 2> gpurun_out/r02_s19_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s19_bench_n2.json").read().strip().splitlines()[-1])
print("N=2", d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["checks"]["gauss_drift_max"], d["checks"]["gauss_ok"], d["checks"]["particles_conserved"])
PY
