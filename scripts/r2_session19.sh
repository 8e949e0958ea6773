#!/bin/bash
# Branch-free edge-id scan in the manifold kernel, no integer modulo in clip: step time, stage times, parity on the GPU.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-extras --no-cpu --no-rays --no-secondary --no-traffic > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s21_bench.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("stages_ms"))
PY
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_capsules_gpu.py -m gpu -x -q 2>&1 | tail -4
