#!/bin/bash
# Round 2, GPU session 4: flex EPA store with in-lane restart (no overflow phase).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_capsules_gpu.py tests/test_bp_persistent.py -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic"
for r in 8 16; do NCB_EPA_REFILL=$r $B > gpurun_out/r2e_bench_r$r.json 2> gpurun_out/r2e_bench.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2e_bench_r*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), {s["stage"]: s["ms"] for s in d["stages_ms"]}, "e2e", round(d["e2e"]["ms_per_step"], 3),
              "restarts", d["counts"].get("n_epa_restarts"), "overflow", d["counts"].get("epa_overflow"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
ncu --set full --clock-control none --import-source on -k regex:"k_cc_epa" -s 6 -c 2 -f -o gpurun_out/r2e_prof_epa \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras --no-traffic > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out | tail -6
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cc_" -s 15 -c 5 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras --no-traffic > gpurun_out/r2e_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r2e_launches.csv | cut -d, -f5,15- | tail -6
