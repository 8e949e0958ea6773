#!/bin/bash
# Round 2, GPU session 8 (N GPUs): routed vs spatial sharding, bench lines.
N=${1:-2}
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic"
for mode in routed spatial; do
  NCB_SHARD=$mode $B > gpurun_out/r2h_n${N}_$mode.json 2> gpurun_out/r2h_n${N}_$mode.err; echo "$mode exit $?"
  tail -c 600 gpurun_out/r2h_n${N}_$mode.err
done
python - <<PY
import json
for mode in ("routed", "spatial"):
    try:
        d = json.loads(open("gpurun_out/r2h_n${N}_%s.json" % mode).read().strip().splitlines()[-1])
        print(mode, round(d["ms_per_step"], 3), {s["stage"]: s["ms"] for s in d["stages_ms"]}, "e2e", round(d["e2e"]["ms_per_step"], 3), "pairs", d["pairs"], "contacts", d["contacts"])
    except Exception as ex:
        print(mode, "ERR", ex)
PY
