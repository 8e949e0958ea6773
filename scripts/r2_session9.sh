#!/bin/bash
# Round 2, GPU session 9 (1 GPU): ray tests (4-wide kernel, uv, per-ray max_toi, pipeline) + ray bench variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rays_r2.py tests/test_gpu_parity.py -x -q -m gpu -k "ray" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
for w in 1 0; do
  NCB_RAY_WIDE=$w timeout 300 python bench.py --rays-only --no-cpu > gpurun_out/r2i_rays_wide$w.json 2> gpurun_out/r2i_rays.err; echo "wide=$w exit $?"
  tail -3 gpurun_out/r2i_rays.err
done
python - <<'PY'
import json
for w in (1, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2i_rays_wide{w}.json").read().strip().splitlines()[-1])
        print("wide", w, "terrain", round(d["ms_per_batch"], 3), "ms", round(d["value"]), "Mrays/s; e2e", round(d["e2e"]["ms_per_batch"], 3), "ms", round(d["e2e"]["value"]), "Mrays/s")
        for k, v in d["variants"].items():
            print("   ", k, round(v["ms_per_batch"], 3), "ms; e2e", round(v["e2e"]["ms_per_batch"], 3))
    except Exception as ex:
        print(w, "ERR", ex)
PY
