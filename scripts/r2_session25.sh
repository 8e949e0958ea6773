#!/bin/bash
# Final line of the round: default bench.py, then the ncu launch list of the same command (short form).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err
tail -c 300 gpurun_out/s25_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s25_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-secondary --no-traffic > gpurun_out/s25_ncu.log 2>&1
tail -2 gpurun_out/s25_ncu.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s25_bench.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "rays", d["rays"]["value"], d["rays"]["e2e"]["value"])
print("dim2", d["widened"]["dim2_contact"]["ms_calls"], d["widened"]["dim2_contact"]["world_update"]["ms"], d["widened"]["dim2_polyline_rays"]["device_ms"])
PY
