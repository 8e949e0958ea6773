#!/bin/bash
# Round 2, GPU session 3: GPU suite incl. capsules / deep tree, bench (overflow EPA beside the manifold kernel), full ncu captures
# of the GJK and manifold kernels.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic"
for v in 1 2 4; do NCB_EPA_BPSM=$v $B > gpurun_out/r2c_bench_over$v.json 2> gpurun_out/r2c_bench.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench_over*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), {s["stage"]: s["ms"] for s in d["stages_ms"]}, "e2e", round(d["e2e"]["ms_per_step"], 3))
    except Exception as ex:
        print(f, "ERR", ex)
PY
ncu --set full --clock-control none --import-source on -k regex:"k_cc_gjk|k_cc_manifold" -s 6 -c 2 -f -o gpurun_out/r2c_prof_gjk_man \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras --no-traffic > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out | tail -8
