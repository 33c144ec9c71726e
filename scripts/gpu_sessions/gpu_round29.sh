#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "subflow or golden or fused or variants" ) 2>&1 | tail -4 | tee gpurun_out/pytest_quick.log
bash scripts/ab_libs.sh 128 libstrugepic_b200_base2.so libstrugepic_b200.so
