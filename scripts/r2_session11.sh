#!/bin/bash
# Round 2, GPU session 11 (1 GPU): what the driver runs at round end: whole GPU suite, smoke, default bench, reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2k_smoke.log
( time python bench.py ) > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"
tail -5 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {s["stage"]: s["ms"] for s in d["stages_ms"]})
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "kernel_ms")})
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("rays", d["rays"]["value"], d["rays"]["e2e"]["value"], d["rays"].get("cpu_baseline", {}).get("value"))
print("secondary", json.dumps(d.get("secondary"))[:600])
print("widened", json.dumps(d.get("widened"))[:800])
PY
