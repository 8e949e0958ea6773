#!/bin/bash
# Round 2, GPU session 2: whole GPU suite with the overflow-stream EPA, then the complete bench line (rays variants, CPU baselines,
# live traffic child) and a launch list.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stages_ms"])
print("roofline", d["roofline"])
print("e2e", d["e2e"])
print("cpu", d["cpu_baseline"])
r = d["rays"]
print("rays", r["value"], r["e2e"], r.get("cpu_baseline"))
for k, v in r["variants"].items():
    print(k, v["value"], v["e2e"]["value"], v["hit_fraction"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras --no-traffic > gpurun_out/r2b_under_ncu.log 2>&1
tail -45 gpurun_out/r2b_launches.csv | cut -d, -f5,12-
