#!/bin/bash
# Round 2, GPU session 13 (1 GPU): occupancy sweeps of the GJK / manifold kernels after the slim-operand change.
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic --no-secondary"
for v in "NCB_GJK_BPSM=4" "NCB_GJK_BPSM=5" "NCB_GJK_BPSM=7" "NCB_MAN_BPSM=3" "NCB_MAN_BPSM=4" "NCB_MAN_BPSM=5" "NCB_EPAS_BPSM=5" "NCB_EPA_REFILL=8" "NCB_EPA_REFILL=24"; do
  env $v $B > gpurun_out/r2m_$v.json 2> gpurun_out/r2m_bench.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2m_{v}.json").read().strip().splitlines()[-1])
    st = {s["stage"]: s["ms"] for s in d["stages_ms"]}
    print(v, round(d["ms_per_step"], 3), "gjk", st["cc_gjk"], "epa", st["cc_epa"], "man", st["cc_manifold"], "e2e", round(d["e2e"]["ms_per_step"], 3))
except Exception as ex:
    print(v, "ERR", ex)
PY
done
