#!/bin/bash
# Eight ranks under torchrun with the late-round build: the driver's scaling launch at N = 8 (sanity + the figure).
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-extras --no-traffic > gpurun_out/s24_n8.json 2> gpurun_out/s24_n8.err
tail -c 400 gpurun_out/s24_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s24_n8.json').read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d.get("sharding"), "pairs", d["pairs"])
print("stages", d["stages_ms"]); print("secondary", {k:v.get("ms_per_update") for k,v in (d.get("secondary") or {}).items()})
PY
