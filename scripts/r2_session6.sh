#!/bin/bash
# Round 2, GPU session 6: whole GPU suite (kinematics, chunked manifold D2H), then the bench with everything.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {s["stage"]: s["ms"] for s in d["stages_ms"]})
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "kernel_ms")})
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print("rays", d["rays"]["value"], d["rays"]["e2e"]["value"], d["rays"]["cpu_baseline"]["value"])
PY
