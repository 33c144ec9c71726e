#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/ab_kernels.py 128 64 0 fusedonly 2>&1 | tee gpurun_out/ab_base.log
SPIC_EXTRA_NVCC_FLAGS="-DSPIC_EXPERIMENT_NO_GENERAL_RED" python -m strugepic_b200.build --force > gpurun_out/build_exp.log 2>&1; tail -2 gpurun_out/build_exp.log
timeout 300 python scripts/ab_kernels.py 128 64 0 fusedonly 2>&1 | tee gpurun_out/ab_nored.log
