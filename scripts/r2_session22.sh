#!/bin/bash
# Validation of the late-round build: whole GPU suite, smoke, default bench line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err
tail -c 600 gpurun_out/s22_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s22_bench.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("traffic"))
print("stages", d["stages_ms"])
print("rays", d["rays"]["value"], d["rays"]["e2e"]["value"])
print("widened keys", list((d.get("widened") or {}).keys()))
print("polyline", (d.get("widened") or {}).get("dim2_polyline_rays"))
print("secondary", d.get("secondary"))
PY
