#!/bin/bash
# round 2, session 2b (2 GPUs): slab parity at world 2 with the one-exchange-per-block protocol (overlapped and serial),
# 2-GPU drivers, N=2 bench overlap on / off
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/r02_s2b_pytest.log
for ov in 1 0; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --opt overlap=$ov > gpurun_out/r02_s2b_bench_n2_ov$ov.json 2> gpurun_out/r02_s2b_bench_n2_ov$ov.err
  tail -c 600 gpurun_out/r02_s2b_bench_n2_ov$ov.err; cat gpurun_out/r02_s2b_bench_n2_ov$ov.json
done
