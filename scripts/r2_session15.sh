#!/bin/bash
# Round 2, GPU session 15: e2e figure, three runs (wall-clock figure: is 5.9 ms noise or a regression?)
mkdir -p gpurun_out
for k in 1 2 3; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic --no-secondary > gpurun_out/r2q_bench_$k.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_bench_$k.json").read().strip().splitlines()[-1])
print($k, round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3))
PY
done
NCB_NO_EARLY_FETCH=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic --no-secondary > gpurun_out/r2q_bench_noearly.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_bench_noearly.json").read().strip().splitlines()[-1])
print("no early fetch", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3))
PY
