#!/bin/bash
# 2 GPUs with the sync-free step path: slab parity at world 2, driver over two ranks, weak-scaling bench line
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_drivers.py -m gpu -x -q -k "2- or two_gpus" ) 2>&1 | tail -6 | tee gpurun_out/pytest_mgpu2_v2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 400 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
