#!/bin/bash
# Shared-tail simplex projections: step time, stage times, parity on the GPU, and one ncu capture of k_cc_gjk.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-extras --no-cpu --no-rays --no-secondary --no-traffic > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s17_bench.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("stages"))
PY
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cc_gjk" -s 6 -c 1 -f -o gpurun_out/s17_prof_gjk python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-rays --no-secondary --no-traffic > gpurun_out/s17_ncu.log 2>&1
tail -3 gpurun_out/s17_ncu.log
