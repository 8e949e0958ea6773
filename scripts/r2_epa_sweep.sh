#!/bin/bash
# Round 2, first GPU session: parity of the shared-memory EPA kernel, then stage times of old vs new and the refill threshold.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-rays"
$B > gpurun_out/r2a_new_r16.json 2> gpurun_out/r2a_new_r16.err
NCB_EPA_SHARED=0 $B > gpurun_out/r2a_old.json 2>/dev/null
for r in 4 8 24 32; do NCB_EPA_REFILL=$r $B > gpurun_out/r2a_new_r$r.json 2>/dev/null; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {s["stage"]: s["ms"] for s in d["stages_ms"]}
        print(f, round(d["ms_per_step"], 3), {k: st[k] for k in ("cc_gjk", "cc_epa", "cc_manifold")}, d["counts"].get("epa_overflow"), d["counts"].get("n_epa_pairs"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
ncu --set full --clock-control none --import-source on -k regex:"k_cc_epa" -s 6 -c 2 -f -o gpurun_out/r2a_prof_epa \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-rays --no-extras > gpurun_out/r2a_ncu.log 2>&1
ls -la gpurun_out | tail -12
