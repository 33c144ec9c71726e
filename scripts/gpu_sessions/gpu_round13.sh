#!/bin/bash
set -x
mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import sys; sys.path[:0]=['.','oracle','tests']
import sys; sys.path.insert(0, "tests/tools"); import debug_fused as d
d.case((7,8,6), 3, 0.15, 0, 2, 3, 1)
d.case((7,8,6), 3, 0.15, 0, 2, 3, 1)
d.case((7,8,6), 3, 0.15, 0, 2, 3, 1, bk=3)
d.case((7,8,6), 3, 0.15, 0, 2, 3, 1, bk=1)
PY
timeout 300 python /tmp/dbg.py 2>&1 | tee gpurun_out/dbg.log
cat > /tmp/dbg2.py <<'PY'
import sys; sys.path[:0]=['.','oracle','tests']
import sys; sys.path.insert(0, "tests/tools"); import debug_fused as d
d.case((7,8,6), 3, 0.15, 0, 2, 1, 1)
PY
timeout 600 compute-sanitizer --tool racecheck python /tmp/dbg2.py 2>&1 | tail -40 | tee gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool initcheck python /tmp/dbg2.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         in" | tail -40 | tee gpurun_out/initcheck.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/dbg2.py 2>&1 | tail -20 | tee gpurun_out/memcheck.log
