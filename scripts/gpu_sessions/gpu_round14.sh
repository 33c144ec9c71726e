#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import sys; sys.path[:0]=['.','oracle','tests']
import sys; sys.path.insert(0, "tests/tools"); import debug_fused as d
d.case((7,8,6), 3, 0.15, 0, 2, 2, 1)
d.case((7,8,6), 3, 0.15, 0, 2, 2, 1, bk=3)
d.case((12,10,7), 40, 0.3, 0, 2, 2, 1)
PY
timeout 300 python /tmp/dbg.py 2>&1 | tee gpurun_out/dbg.log
