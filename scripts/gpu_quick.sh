#!/bin/bash
# quick loop: fused parity tests + A/B timing of the fused kernels
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or golden or crossings or ragged" ) 2>&1 | tail -8 | tee gpurun_out/pytest_quick.log
timeout 300 python scripts/ab_kernels.py 128 64 0 fusedonly 2>&1 | tee gpurun_out/ab_quick.log
