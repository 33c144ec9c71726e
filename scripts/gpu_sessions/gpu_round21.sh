#!/bin/bash
# 4 GPUs: slab-decomposition parity at world = 2 and 4, short weak-scaling bench at N = 4
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) 2>&1 | tail -30 | tee gpurun_out/pytest_mgpu4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -c 800 gpurun_out/bench_n4.err; cat gpurun_out/bench_n4.json
