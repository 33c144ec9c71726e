#!/bin/bash
mkdir -p gpurun_out
for lib in libstrugepic_b200.so libstrugepic_b200_p4b2.so; do echo "== $lib"; SPIC_B200_LIBRARY=$PWD/strugepic_b200/lib/$lib timeout 100 python tests/tools/check_push4.py 2>&1 | tail -12; done | tee gpurun_out/check_push4.log
