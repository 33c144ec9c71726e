#!/bin/bash
# A/B: conflict-free deposition record layout (SW = 12 + group swizzle) against the previous build, plus fused parity
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or golden or crossings or ragged or deferred" ) 2>&1 | tail -5 | tee gpurun_out/pytest_quick.log
bash scripts/ab_libs.sh 128 libstrugepic_b200_base.so libstrugepic_b200.so libstrugepic_b200_base.so libstrugepic_b200.so
