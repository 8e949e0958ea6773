#!/bin/bash
# Round 2, GPU session 8 (N GPUs): p2p vs routed (vs spatial) sharding, bench lines, topology.
N=${1:-2}
MODES=${2:-"p2p routed"}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo_n$N.txt 2>&1
(ls /sys/devices/system/node/ | grep node; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null) >> gpurun_out/r2h_topo_n$N.txt 2>&1
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-extras --no-rays --no-traffic $EXTRA"
for mode in $MODES; do
  NCB_SHARD=$mode timeout 300 $B > gpurun_out/r2h_n${N}_$mode.json 2> gpurun_out/r2h_n${N}_$mode.err; echo "$mode exit $?"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/r2h_n${N}_$mode.err | tail -5
done
python - <<PY
import json
for mode in "$MODES".split():
    try:
        d = json.loads(open("gpurun_out/r2h_n${N}_%s.json" % mode).read().strip().splitlines()[-1])
        print(mode, round(d["ms_per_step"], 3), {s["stage"]: s["ms"] for s in d["stages_ms"]}, "e2e", round(d["e2e"]["ms_per_step"], 3), "pairs", d["pairs"], "contacts", d["contacts"], d.get("sharding"))
    except Exception as ex:
        print(mode, "ERR", ex)
PY
