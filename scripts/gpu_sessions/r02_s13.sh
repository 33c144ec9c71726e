#!/bin/bash
# round 2, session 13 (2 GPUs): the driver's scaling command at N=2, e2e leg included
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_s13_bench_n2.json 2> gpurun_out/r02_s13_bench_n2.err
tail -c 800 gpurun_out/r02_s13_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s13_bench_n2.json").read().strip().splitlines()[-1])
print("N=2", d["value"], d["ms_per_step"], d["e2e"], d["checks"]["gauss_ok"], d["checks"]["particles_conserved"], d["gpu_launches"])
PY
