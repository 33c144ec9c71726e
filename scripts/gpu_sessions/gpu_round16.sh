#!/bin/bash
# 2 GPUs: slab-decomposition parity (fused and per-sub-flow schedules) + a short weak-scaling bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) 2>&1 | tail -30 | tee gpurun_out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1200 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
