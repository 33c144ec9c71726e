#!/bin/bash
# round 2, session 5 (1 GPU): half-blocks on wall boxes, double-buffered get_particles, whole suite, full bench line
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/r02_s5_pytest.log 2>&1
grep -E "passed|failed|FAILED|PARITY|rror" gpurun_out/r02_s5_pytest.log | head -40
timeout 900 python bench.py > gpurun_out/r02_s5_bench.json 2> gpurun_out/r02_s5_bench.err
tail -c 600 gpurun_out/r02_s5_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s5_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
for s in d["secondary"]: print(s.get("name"), s.get("value"), s.get("ms_per_step"), s.get("error"))
PY
