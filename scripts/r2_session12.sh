#!/bin/bash
# Round 2, GPU session 12 (1 GPU): slim GJK operands: parity + bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_capsules_gpu.py tests/test_bp_persistent.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic --no-secondary"
for bpsm in 6 8; do
NCB_GJK_BPSM=$bpsm $B > gpurun_out/r2l_bench_$bpsm.json 2> gpurun_out/r2l_bench.err; echo "exit $?"
done
python - <<'PY'
import json
for f in (6, 8):
    try:
        d = json.loads(open(f"gpurun_out/r2l_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), {s["stage"]: s["ms"] for s in d["stages_ms"]}, "e2e", round(d["e2e"]["ms_per_step"], 3))
    except Exception as ex:
        print(f, "ERR", ex)
PY
