#!/bin/bash
# A/B of library builds on the fused path: scripts/ab_libs.sh <cells> <lib> [<lib> ...]  (lib = file name under strugepic_b200/lib)
cells=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  SPIC_B200_LIBRARY=$PWD/strugepic_b200/lib/$lib timeout 300 python scripts/ab_kernels.py $cells 64 0 fusedonly 2>&1 | grep -E "fused|energy"
done 2>&1 | tee -a gpurun_out/ab_libs.log
