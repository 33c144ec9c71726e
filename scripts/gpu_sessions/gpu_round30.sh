#!/bin/bash
# 2 GPUs: the driver binary over two ranks (file rendezvous) and the slab parity tests at world 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_drivers.py -m gpu -x -q -k "two_gpus" ) 2>&1 | tail -15 | tee gpurun_out/pytest_driver_2gpu.log
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "2-" ) 2>&1 | tail -8 | tee gpurun_out/pytest_mgpu2.log
